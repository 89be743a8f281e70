"""Test data for ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1222-1302) — TEST INFRASTRUCTURE ONLY (uses the oracle to place the
true correspondences).  Two keyframes looking at the same map points from nearby poses, related by the similarity (s12 = 1, R12, t12)
the loop closer would have estimated; each keypoint may hold a map point of its own keyframe."""
import numpy as np

import oracle_lib as O
from vieo_slam_b200 import synth
from vieo_slam_b200.layouts import KP_DTYPE, PROJ_SEARCH_FRAME_DTYPE

TH_HIGH = 100


def make(seed, th=7.5, n_kp=1200, n_q=1400, prior_frac=0.05):
    """-> (side1, side2, sim3, prior12): side = dict(frame, kps, uright, desc, wP, Pn, maxd, mind, qdesc, skip) of one keyframe with its
    OWN map point per keypoint (skip = 1: none); sim3 = (s12, R12 [9] f32, t12 [3] f32)."""
    r = np.random.default_rng(seed + 500)
    pb = synth.make_fuse_problem(seed, n_frames=1, n_kp=n_kp, n_q=n_q, th_radius=th, use_bf=False, check_viewing_angle=False, skip_frac=0.1)
    F2 = pb["frames"][0].copy()
    N2, N1 = int(F2["n_kp"]), int(F2["n_q"])
    best, dist, lvl = O.proj_search(pb)
    # keyframe 2: keypoints of the problem; its map point at keypoint idx2 = the first query found there (85 % of them)
    skip2 = np.ones(N2, np.uint8); src2 = np.full(N2, -1)
    for i1 in range(N1):
        j = best[i1]
        if j >= 0 and dist[i1] <= TH_HIGH and src2[j] < 0:
            src2[j] = i1
    has2 = (src2 >= 0) & (r.random(N2) < 0.85)
    skip2[has2] = 0
    s2 = np.where(has2, src2, 0)
    side2 = dict(frame=F2, kps=pb["kps"][:N2].copy(), uright=pb["uright"][:N2].copy(), desc=pb["desc"][:N2].copy(),
                 wP=pb["p_wP"][s2].copy(), Pn=pb["p_normal"][s2].copy(), maxd=pb["p_max_dist"][s2].copy(), mind=pb["p_min_dist"][s2].copy(),
                 qdesc=pb["desc"][:N2].copy(), skip=skip2)
    # keyframe 1: a nearby pose; keypoint i1 = projection of query i1 (its own map point)
    R2 = np.asarray(F2["Rcw"], np.float64).reshape(3, 3); t2 = np.asarray(F2["tcw"], np.float64)
    R1 = (synth.so3_exp(r.normal(0, 0.01, 3)) @ R2).astype(np.float32); t1 = (t2 + r.normal(0, 0.03, 3)).astype(np.float32)
    O1 = (-(R1.astype(np.float64).T @ t1.astype(np.float64))).astype(np.float32)
    Pc = (R1.astype(np.float64) @ pb["p_wP"][:N1].astype(np.float64).T).T + t1.astype(np.float64)
    z = np.where(Pc[:, 2] > 0.05, Pc[:, 2], 1.0)
    u = F2["fx"] * Pc[:, 0] / z + F2["cx"]; v = F2["fy"] * Pc[:, 1] / z + F2["cy"]
    inside = (Pc[:, 2] > 0.05) & (u > F2["minx"] + 2) & (u < F2["maxx"] - 2) & (v > F2["miny"] + 2) & (v < F2["maxy"] - 2)
    skip1 = ((pb["p_skip"][:N1] != 0) | ~inside).astype(np.uint8)
    k1 = np.zeros(N1, KP_DTYPE)
    k1["x"] = np.where(inside, u + r.normal(0, 0.4, N1), r.uniform(F2["minx"] + 5, F2["maxx"] - 5, N1)).astype(np.float32)
    k1["y"] = np.where(inside, v + r.normal(0, 0.4, N1), r.uniform(F2["miny"] + 5, F2["maxy"] - 5, N1)).astype(np.float32)
    k1["octave"] = np.where(lvl[:N1] >= 0, lvl[:N1], r.integers(0, 8, N1)); k1["size"] = 31
    d1 = pb["q_desc"][:N1].copy()
    for i in range(N1):                                                  # a few flipped bits: the keypoint is not the map point's copy
        bits = r.choice(256, int(r.integers(0, 12)), replace=False)
        np.bitwise_xor.at(d1[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    F1 = F2.copy()
    F1["Rcw"] = R1.reshape(-1); F1["tcw"] = t1; F1["Ow"] = O1; F1["n_kp"] = N1; F1["n_q"] = N1
    side1 = dict(frame=F1, kps=k1, uright=np.full(N1, -1, np.float32), desc=d1, wP=pb["p_wP"][:N1].copy(), Pn=pb["p_normal"][:N1].copy(),
                 maxd=pb["p_max_dist"][:N1].copy(), mind=pb["p_min_dist"][:N1].copy(), qdesc=pb["q_desc"][:N1].copy(), skip=skip1)
    F2["n_q"] = N2
    R12 = (R1.astype(np.float64) @ R2.T).astype(np.float32)
    t12 = (t1.astype(np.float64) - R12.astype(np.float64) @ t2).astype(np.float32)
    prior12 = np.full(N1, -1, np.int32)
    cand = np.nonzero((best[:N1] >= 0) & (skip1 == 0) & (skip2[np.maximum(best[:N1], 0)] == 0))[0]
    pick = cand[r.random(len(cand)) < prior_frac]
    used = set()
    for i1 in pick:
        if int(best[i1]) not in used:
            prior12[i1] = best[i1]; used.add(int(best[i1]))
    return side1, side2, (np.float32(1.0), R12.reshape(-1), t12), prior12


def flat_problem(side1, side2, pose21, pose12, prior12, th):
    """The two searches of SearchBySim3 as ONE two-frame SearchByProjectionBase batch (make_fuse_problem layout): frame 0 = keyframe 1's
    points into keyframe 2 with pose (sR21 R1w, sR21 t1w + t21), frame 1 = keyframe 2's points into keyframe 1 with (sR12 R2w, sR12 t2w
    + t12); already-matched points are skipped (vbAlreadyMatched1 / 2)."""
    N1, N2 = len(side1["kps"]), len(side2["kps"])
    fr = np.zeros(2, PROJ_SEARCH_FRAME_DTYPE)
    fr[0] = side2["frame"]; fr[1] = side1["frame"]
    fr[0]["kp_begin"], fr[0]["n_kp"], fr[0]["q_begin"], fr[0]["n_q"] = 0, N2, 0, N1
    fr[1]["kp_begin"], fr[1]["n_kp"], fr[1]["q_begin"], fr[1]["n_q"] = N2, N1, N1, N2
    fr[0]["Rcw"], fr[0]["tcw"] = pose21[:9], pose21[9:12]
    fr[1]["Rcw"], fr[1]["tcw"] = pose12[:9], pose12[9:12]
    fr["th_radius"] = th; fr["use_bf"] = 0; fr["check_viewing_angle"] = 0
    skipA = side1["skip"].copy(); skipA[prior12 >= 0] = 1
    skipB = side2["skip"].copy(); skipB[prior12[prior12 >= 0]] = 1
    cat = lambda k: np.ascontiguousarray(np.concatenate([side2[k], side1[k]]))      # keypoint arrays: keyframe 2 first
    catq = lambda k: np.ascontiguousarray(np.concatenate([side1[k], side2[k]]))     # query arrays: keyframe 1's points first
    return dict(frames=fr, kps=cat("kps"), uright=cat("uright"), desc=cat("desc"), p_wP=catq("wP"), p_normal=catq("Pn"),
                p_max_dist=catq("maxd"), p_min_dist=catq("mind"), q_desc=catq("qdesc"), p_skip=np.concatenate([skipA, skipB]).astype(np.uint8),
                has_mp=np.concatenate([side2["skip"] == 0, side1["skip"] == 0]), prior12=prior12.copy())
