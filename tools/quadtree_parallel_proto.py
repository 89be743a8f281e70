#!/usr/bin/env python3
"""Prototype of the data-parallel formulation of DistributeOctTree used by the CUDA kernel
(vieo_slam_b200/csrc/orb_quadtree.cu), checked against the sequential oracle.  Every step is a
"for all keys" / "for all nodes" map plus prefix sums — no linked list.  Dev tool, not shipped."""
import math, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))


def quadtree_parallel(xyr, W, H, N):
    B = 16
    x = xyr[:, 0].astype(np.float32) - B
    y = xyr[:, 1].astype(np.float32) - B
    resp = xyr[:, 2]
    nk = len(xyr)
    minX, maxX, minY, maxY = B, W - 16, B, H - 16
    f32 = np.float32
    nIni = int(np.round(f32(maxX - minX) / f32(maxY - minY)))  # C round(): half away; fine for these ratios
    hX = f32(maxX - minX) / f32(nIni)
    # node table in LIST ORDER
    nodes = []  # dict(x0,x1,y0,y1,cnt,single,id)
    key_node = (x / hX).astype(np.int32)
    cnt = np.bincount(key_node, minlength=nIni)
    remap = -np.ones(nIni, np.int32)
    for i in range(nIni):
        if cnt[i] == 0:
            continue
        remap[i] = len(nodes)
        nodes.append(dict(x0=int(hX * f32(i)), x1=int(hX * f32(i + 1)), y0=0, y1=maxY - minY, cnt=int(cnt[i]),
                          single=cnt[i] == 1, id=i))
    key_node = remap[key_node]
    next_id = nIni

    def child_counts(sel_nodes):
        """for nodes in sel (list positions) -> mids and per-key quadrant, counts[node][4]"""
        mx = np.zeros(len(nodes), np.int32); my = np.zeros(len(nodes), np.int32)
        for p in sel_nodes:
            n = nodes[p]
            mx[p] = n["x0"] + math.ceil(f32(n["x1"] - n["x0"]) / 2)
            my[p] = n["y0"] + math.ceil(f32(n["y1"] - n["y0"]) / 2)
        sel = np.zeros(len(nodes), bool); sel[sel_nodes] = True
        ksel = sel[key_node]
        q = (x >= mx[key_node]).astype(np.int32) + 2 * (y >= my[key_node]).astype(np.int32)
        counts = np.zeros((len(nodes), 4), np.int32)
        np.add.at(counts, (key_node[ksel], q[ksel]), 1)
        return mx, my, q, ksel, counts

    def apply(processed, mx, my, q, ksel, counts):
        """processed: list positions in processing order. Build new list: reversed children + untouched."""
        nonlocal nodes, key_node, next_id
        children = []
        newpos_of = {}
        for p in processed:
            n = nodes[p]
            bounds = [(n["x0"], mx[p], n["y0"], my[p]), (mx[p], n["x1"], n["y0"], my[p]),
                      (n["x0"], mx[p], my[p], n["y1"]), (mx[p], n["x1"], my[p], n["y1"])]
            for qq in range(4):
                c = int(counts[p, qq])
                if c == 0:
                    continue
                children.append(dict(x0=bounds[qq][0], x1=bounds[qq][1], y0=bounds[qq][2], y1=bounds[qq][3], cnt=c,
                                     single=c == 1, id=next_id + len(children)))
                newpos_of[(p, qq)] = len(children) - 1
        C = len(children)
        pset = set(processed)
        keep = [i for i in range(len(nodes)) if i not in pset]
        keep_pos = {old: C + r for r, old in enumerate(keep)}
        new_nodes = children[::-1] + [nodes[i] for i in keep]
        new_key_node = key_node.copy()
        for k in range(nk):
            p = key_node[k]
            if p in pset:
                new_key_node[k] = C - 1 - newpos_of[(p, int(q[k]))]
            else:
                new_key_node[k] = keep_pos[p]
        new_ids = [c["id"] for c in children if c["cnt"] > 1]
        next_id += C
        nodes = new_nodes
        key_node = new_key_node
        return new_ids

    finish = False
    while not finish:
        prev = len(nodes)
        todo = [i for i, n in enumerate(nodes) if not n["single"]]
        mx, my, q, ksel, counts = child_counts(todo)
        exp_ids = apply(todo, mx, my, q, ksel, counts)
        if len(nodes) >= N or len(nodes) == prev:
            finish = True
        elif len(nodes) + 3 * len(exp_ids) > N:
            while not finish:
                prev2 = len(nodes)
                idpos = {n["id"]: i for i, n in enumerate(nodes)}
                E = sorted(((nodes[idpos[i]]["cnt"], i) for i in exp_ids), reverse=True)
                pos = [idpos[i] for _, i in E]
                mx, my, q, ksel, counts = child_counts(pos)
                delta = [(counts[p] > 0).sum() - 1 for p in pos]
                run = prev2
                take = len(pos)
                for j, d in enumerate(delta):
                    run += d
                    if run >= N:
                        take = j + 1
                        break
                exp_ids = apply(pos[:take], mx, my, q, ksel, counts)
                if len(nodes) >= N or len(nodes) == prev2:
                    finish = True
    picked = []
    for i, n in enumerate(nodes):
        ks = np.nonzero(key_node == i)[0]
        assert len(ks) == n["cnt"]
        picked.append(int(ks[np.argmax(resp[ks])]))  # argmax -> first max
    return np.array(picked, np.int32)


if __name__ == "__main__":
    import oracle_lib as O
    rng = np.random.default_rng(0)
    for trial in range(40):
        W, H = [(752, 480), (627, 400), (210, 134), (512, 512), (252, 161)][trial % 5]
        n = int(rng.choice([0, 1, 2, 5, 50, 300, 2000, 6000]))
        pts = set()
        while len(pts) < n:
            pts.add((int(rng.integers(19, W - 19)), int(rng.integers(19, H - 19))))
        pts = sorted(pts, key=lambda p: (p[1] // 38, p[0] // 36, p[1], p[0]))
        xyr = np.array([(px, py, int(rng.integers(7, 60))) for px, py in pts], np.int32).reshape(-1, 3)
        for N in (72, 261, 1000):
            a = O.quadtree(xyr, W, H, N)
            b = quadtree_parallel(xyr, W, H, N) if n else np.zeros(0, np.int32)
            assert np.array_equal(a, b), (trial, W, H, n, N, len(a), len(b))
    print("parallel formulation == sequential oracle on all trials")
