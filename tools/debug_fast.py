import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, oracle_lib as O
import vieo_slam_b200.api as api
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cv2_orb_goldens.npz"))
img = g["A_img"]
orb = api.ORBextractor(1200, 1.2, 8, 20, 7, 752, 480)
ret, kps, desc = orb(img, want_pyramid=True)
c = orb.debug_candidates(0, 0)
S = O.fast_score_map(img)
print("gpu cands", len(c), c[:12].tolist())
print("true S at gpu cands:", [int(S[y, x]) for x, y, r in c[:12]])
# is response maybe score of a different pixel? search neighbourhood
for x, y, r in c[:5]:
    print((x, y, r), "S patch", S[y-3:y+4, x-3:x+4].tolist())
