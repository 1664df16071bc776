"""Host-side packing of weight matrices into the shared-memory images the tcgen05 kernels read.

`pack_b_sw128(w[n][k])` -> uint8 image: K is cut into slices of 64 fp16 elements (one 128-byte swizzle row per
matrix row and slice); slice `s` is an [n x 128 B] block at byte offset s * n * 128, and inside a row the 16-byte
chunk `c` (8 elements) sits at chunk position c ^ (row & 7) -- the K-major SWIZZLE_128B layout of the UMMA shared
memory descriptor (csrc/tc_common.cuh: smem_desc_sw128).  Kernels move the image with one cp.async.bulk.
"""
import numpy as np


def pack_b_sw128(w):
    w = np.asarray(w, np.float32)
    n, k = w.shape
    nsl = -(-k // 64)
    wp = np.zeros((n, nsl * 64), np.float16)
    wp[:, :k] = w.astype(np.float16)
    img = np.zeros((nsl, n, 8, 8), np.float16)          # [slice][row][chunk position][8 elements]
    rows = np.arange(n)
    for s in range(nsl):
        blk = wp[:, s * 64:(s + 1) * 64].reshape(n, 8, 8)
        for c in range(8):
            img[s, rows, c ^ (rows & 7)] = blk[:, c]
    return img.reshape(-1).view(np.uint8)


def pack_b_sw64(w):
    """[n][k] -> K slices of 32 fp16 (64-byte rows), slice s an [n x 64 B] block at byte offset s * n * 64; 16-byte
    chunk c of row r sits at chunk position c ^ ((r >> 1) & 3) -- the K-major SWIZZLE_64B layout (smem_desc_sw64)."""
    w = np.asarray(w, np.float32)
    n, k = w.shape
    nsl = -(-k // 32)
    wp = np.zeros((n, nsl * 32), np.float16)
    wp[:, :k] = w.astype(np.float16)
    img = np.zeros((nsl, n, 4, 8), np.float16)
    rows = np.arange(n)
    for s in range(nsl):
        blk = wp[:, s * 32:(s + 1) * 32].reshape(n, 4, 8)
        for c in range(4):
            img[s, rows, c ^ ((rows >> 1) & 3)] = blk[:, c]
    return img.reshape(-1).view(np.uint8)
