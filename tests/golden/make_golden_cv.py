"""Generates tests/golden/newpts_cv.npz with OpenCV itself (cv2 4.x, Python wheel): the occupancy
mask of DefLocalMapping::CreateNewMapPoints / needNewTemplate (DefLocalMapping.cc:240-347,355-403)
exactly as the reference computes it -- paint, cv2.filter2D with a ones kernel of edge cols/20,
cv2.threshold(1, 255, THRESH_BINARY) -- read back at the keypoint pixels, and the fp32 cv::Mat
product Twc * x3ch through cv2.gemm.  A true pin of the oracle (and of the kernel) against the
library the reference calls.

    python tests/golden/make_golden_cv.py
"""
import os

import cv2
import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))


def make_case(seed, n, rows, cols, border_heavy=False):
    rng = np.random.default_rng(seed)
    xy = np.stack([rng.uniform(0, cols - 0.01, n), rng.uniform(0, rows - 0.01, n)], 1).astype(np.float32)
    if border_heavy:  # keypoints hugging the image borders exercise BORDER_REFLECT_101
        k = n // 2
        xy[:k, 0] = rng.choice([0.2, 1.7, 3.0, cols - 1.2, cols - 2.5, cols - 0.6], k).astype(np.float32)
        xy[k // 2:k, 1] = rng.choice([0.4, 2.2, rows - 1.1, rows - 3.4], k - k // 2).astype(np.float32)
    state = rng.choice([0, 1, 2], n, p=[0.55, 0.35, 0.10]).astype(np.uint8)
    # cluster the mapped points so that part of the image is unoccupied
    m = state == 1
    xy[m, 0] = np.clip(xy[m, 0] * 0.55 + (0.0 if seed % 2 else cols * 0.4), 0, cols - 0.01)
    surf = rng.normal(size=(n, 3)).astype(np.float32) * 0.3 + np.float32([0, 0, 1.0])
    ang = rng.normal(size=3) * 0.2
    R, _ = cv2.Rodrigues(ang)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = R.astype(np.float32)
    T[:3, 3] = (rng.normal(size=3) * 0.1).astype(np.float32)
    return xy, state, surf, T


def reference_cv(xy, state, surf, T, rows, cols):
    n = len(state)
    mask = np.zeros((rows, cols), np.uint8)
    for i in range(n):
        if state[i] == 1:
            mask[int(xy[i, 1]), int(xy[i, 0])] = 255
    ksz = cols // 20
    kernel = np.ones((ksz, ksz), np.float32)
    mask = cv2.filter2D(mask, -1, kernel, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_DEFAULT)
    _, mask = cv2.threshold(mask, 1, 255, 0)
    action = np.zeros(n, np.uint8)
    world = np.zeros((n, 3), np.float32)
    for i in range(n):
        if state[i] == 1:
            action[i] = 1
        elif state[i] == 0 and not mask[int(xy[i, 1]), int(xy[i, 0])]:
            action[i] = 2
        if action[i]:
            x3ch = np.float32([[surf[i, 0]], [surf[i, 1]], [surf[i, 2]], [1.0]])
            x3wh = cv2.gemm(T, x3ch, 1.0, None, 0.0)
            world[i] = x3wh[:3, 0]
    return action, world, int((action == 2).sum())


def main():
    out = {}
    cases = [(11, 1200, 480, 640, False), (12, 900, 480, 640, True), (13, 300, 288, 360, True), (14, 40, 60, 45, True)]
    out["cases"] = np.array(cases, dtype=np.int64)
    for ci, (seed, n, rows, cols, bh) in enumerate(cases):
        xy, state, surf, T = make_case(seed, n, rows, cols, bh)
        action, world, n_new = reference_cv(xy, state, surf, T, rows, cols)
        out[f"xy{ci}"], out[f"state{ci}"], out[f"surf{ci}"], out[f"T{ci}"] = xy, state, surf, T
        out[f"action{ci}"], out[f"world{ci}"], out[f"nnew{ci}"] = action, world, np.int64(n_new)
        print(ci, n, rows, cols, "new", n_new, "moved", int((action == 1).sum()))
    out["cv_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "newpts_cv.npz"), **out)


if __name__ == "__main__":
    main()
