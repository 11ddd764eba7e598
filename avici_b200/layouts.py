"""Index maps of the tensor-core path's activation layouts (host-side mirror of `RowMap` and the operand layouts in
csrc/kernels_tc.cuh).  The kernels never call this: it documents the layouts, lets the CPU test
suite check that they are bijections onto the buffers the workspace reserves, and is what a debugger needs to pull a
token's q / k / v or attention row out of a workspace dump."""
import numpy as np


def seq_major_to_token(m, n, d, axis):
    """Sequence-major row index -> token index t = (b*n + i)*d + j (RowMap::token).
    axis 0 (attend over d): sequence (b, i), position j: m == t.  axis 1 (attend over n): sequence (b, j), position i."""
    m = np.asarray(m, dtype=np.int64)
    if axis == 0:
        return m
    b, rem = np.divmod(m, n * d)
    j, i = np.divmod(rem, n)
    return (b * n + i) * d + j


def token_to_seq_major(t, n, d, axis):
    t = np.asarray(t, dtype=np.int64)
    if axis == 0:
        return t
    b, rem = np.divmod(t, n * d)
    i, j = np.divmod(rem, d)
    return (b * d + j) * n + i


def qkv_operand_offset(m, which, head, ch, n, d, axis, heads=8):
    """Element offset of channel `ch` (0..31) of head `head` of q/k/v (`which` 0/1/2) of sequence-major row m in the
    attention operand layout qkv[seq][which][head][pos][32], 16-byte chunks stored at chunk ^ ((pos >> 1) & 3)."""
    T = d if axis == 0 else n
    m = np.asarray(m, dtype=np.int64)
    seq, pos = np.divmod(m, T)
    chunk = (ch // 8) ^ ((pos >> 1) & 3)
    return (((seq * 3 + which) * heads + head) * T + pos) * 32 + chunk * 8 + ch % 8


def attn_operand_offset(m, head, ch, heads=8):
    """Element offset in the out-projection's A-operand layout attn[m >> 7][head][m & 127][32], same chunk swizzle."""
    m = np.asarray(m, dtype=np.int64)
    chunk = (ch // 8) ^ ((m >> 1) & 3)
    return (((m >> 7) * heads + head) * 128 + (m & 127)) * 32 + chunk * 8 + ch % 8
