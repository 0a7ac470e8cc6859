import sys, os, itertools, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_umma import *
KC = 32
rng = np.random.default_rng(0)
H = bf16_round(rng.normal(size=(KC, 64))).float().numpy()
L = bf16_round(rng.normal(size=(KC, 64)) * 2.0 ** -9).float().numpy()
R = np.zeros((KC, 64), np.float32); R[:, :16] = bf16_round(rng.integers(1, 6, (KC, 16))).float().numpy()
img = mn_major_image([H, L, R], KC); blk = KC * 128
A = np.concatenate([H, L], 1); B = np.concatenate([H, R[:, :16]], 1)
want = A.T.astype(np.float64) @ B.astype(np.float64)
for name, (al, asb, bl, bsb) in {"lbo=atomMN,sbo=1024": (blk, 1024, 2 * blk, 1024), "swapped": (1024, blk, 1024, 2 * blk)}.items():
    for nm in (1, 2):
        D = probe(img, 0, 0, al, asb, bl, bsb, 2, idesc(1, True, True, 128, 80), nm, 2048, 2048, False, 80)
        w = (A[:16 * nm].T.astype(np.float64) @ B[:16 * nm].astype(np.float64))
        print("MN-major", name, "n_mma", nm, "maxerr", np.abs(D - w).max(), "top-left ok", np.abs(D[:64, :64] - w[:64, :64]).max(),
              "LtH ok", np.abs(D[64:, :64] - w[64:, :64]).max(), "R ok", np.abs(D[:, 64:] - w[:, 64:]).max())
for N in (64, 256):
    A2 = bf16_round(rng.normal(size=(128, 64))).float().numpy(); B2 = bf16_round(rng.normal(size=(N, 64))).float().numpy()
    img2 = np.concatenate([k_major_image(A2, 128), k_major_image(B2, N)])
    for lbo in (16, 1024):
        for nm in (1, 4):
            D = probe(img2, 0, 128 * 128, lbo, 1024, lbo, 1024, 2, idesc(1, False, False, 128, N), nm, 32, 32, False, N)
            w = A2[:, :16 * nm].astype(np.float64) @ B2[:, :16 * nm].astype(np.float64).T
            print("K-major N", N, "lbo", lbo, "n_mma", nm, "maxerr", np.abs(D - w).max())
