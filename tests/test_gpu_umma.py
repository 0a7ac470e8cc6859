"""GPU: pins the tcgen05 (UMMA) shared-memory operand layouts used by the tensor-core kernels
against a numpy product, through the hals_debug_umma_probe self-test entry."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nat():
    import hybrid_als_twotower_recommender_b200  # noqa: F401
    from hybrid_als_twotower_recommender_b200 import _native
    return _native


def bf16_round(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16)


def bf16_bits(x):
    return bf16_round(x).view(torch.int16).numpy().view(np.uint16)


def idesc(fmt, a_mn, b_mn, M, N):
    return (1 << 4) | (fmt << 7) | (fmt << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def swz128(row, byte_in_row):
    """128B swizzle: the 16-byte chunk index is XORed with (row mod 8)."""
    chunk, within = byte_in_row >> 4, byte_in_row & 15
    return ((chunk ^ (row & 7)) << 4) | within


def probe(img, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, swizzle, idsc, n_mma, a_step, b_step, tf32, ncols):
    nat = _nat()
    d_img = torch.from_numpy(img).cuda()
    out = torch.zeros((128, ncols), dtype=torch.float32, device="cuda")
    nat.check(nat.lib().hals_debug_umma_probe(nat.ptr(d_img), img.size, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo,
                                              swizzle, idsc, n_mma, a_step, b_step, int(tf32), ncols, nat.ptr(out),
                                              nat.current_stream()), "probe")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def mn_major_image(blocks, KC):
    """blocks: list of [KC, 64] bf16-representable arrays (one 64-wide MN atom each, rows = K index).
    Each block is KC rows of 128 bytes, swizzled in groups of 8 rows; blocks are laid end to end."""
    img = np.zeros(len(blocks) * KC * 128, np.uint8)
    v16 = img.view(np.uint16)
    for b, blk in enumerate(blocks):
        bits = bf16_bits(blk)
        for t in range(KC):
            for f in range(64):
                off = b * KC * 128 + t * 128 + swz128(t, f * 2)
                v16[off >> 1] = bits[t, f]
    return img


@pytest.mark.parametrize("KC", [16, 64])
def test_mn_major_sw128_bf16_als_layout(KC):
    """ALS rank-64 tile: A = [H ; L] (M = 128, MN-major), B = [H ; R] (N = 80), K = ratings."""
    rng = np.random.default_rng(KC)
    H = bf16_round(rng.normal(size=(KC, 64))).float().numpy()
    L = bf16_round(rng.normal(size=(KC, 64)) * 2.0 ** -9).float().numpy()
    R = np.zeros((KC, 64), np.float32); R[:, :16] = bf16_round(rng.integers(1, 6, (KC, 16))).float().numpy()
    img = mn_major_image([H, L, R], KC)
    blk = KC * 128
    D = probe(img, 0, 0, blk, 1024, 2 * blk, 1024, 2, idesc(1, True, True, 128, 80), KC // 16, 2048, 2048, False, 80)
    A = np.concatenate([H, L], 1)                      # [KC, 128]
    B = np.concatenate([H, R[:, :16]], 1)              # [KC, 80]
    want = A.T.astype(np.float64) @ B.astype(np.float64)
    assert np.allclose(D, want, rtol=1e-5, atol=1e-4), np.abs(D - want).max()


def k_major_image(Mat, rows_pad):
    """Mat [rows, 64] bf16 (one 64-element = 128-byte K block): row r at r*128, swizzled by r mod 8."""
    img = np.zeros(rows_pad * 128, np.uint8)
    v16 = img.view(np.uint16)
    bits = bf16_bits(Mat)
    for r in range(Mat.shape[0]):
        for k in range(64):
            v16[(r * 128 + swz128(r, k * 2)) >> 1] = bits[r, k]
    return img


@pytest.mark.parametrize("N", [64, 128, 256])
def test_k_major_sw128_bf16_scoring_layout(N):
    """Scoring tile: A = users [128, K], B = items [N, K], both K-major, one 64-wide K block = 4 MMAs."""
    rng = np.random.default_rng(N)
    A = bf16_round(rng.normal(size=(128, 64))).float().numpy()
    B = bf16_round(rng.normal(size=(N, 64))).float().numpy()
    img = np.concatenate([k_major_image(A, 128), k_major_image(B, N)])
    D = probe(img, 0, 128 * 128, 16, 1024, 16, 1024, 2, idesc(1, False, False, 128, N), 4, 32, 32, False, N)
    want = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.allclose(D, want, rtol=1e-5, atol=1e-4), np.abs(D - want).max()
