#!/usr/bin/env python
"""CPU emulation of the index arithmetic of the experimental fp16 + e4m3 modes (options "fc_f16f8", "conv_f16f8"):
    pack_b_f16f8_kernel, pack_conv_f16f8_kernel, split16_f16f8, the tape writers (tapgemm EPI_FC_TAPE with F8,
    block2_kernel<true>, block1_kernel<.., 1>), the F8 producer / issuer of tapgemm_kernel and the slabs, weight
    blocks and issuer of block2_kernel<true, true> (csrc/dce_tc.cuh, dce_tc_block1.cuh, dce_tc_block2.cuh)
restated in numpy: bulk copies land in a byte array shaped like a ring stage, every MMA gathers its operands
through the (start, LBO, SBO = 128) descriptor it would be issued with, corrections accumulate at 2^15 and the
first fp16 MMA applies scale-input-d.  Checked against float64 x @ W^T.  No GPU: this pins the layout formulas and
the operand pairing, not the synchronisation.       python tools/emulate_f16f8.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GUARD, SLAB = 8, 130 * 16
EXL, EWH, EXH, EWL, SCALE_D = 12, 3, 1, 14, 15


def e4m3(v):
    """float array -> e4m3 bytes (round to nearest even, saturating, as cvt.rn.satfinite.e4m3x2.f32)."""
    t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).clamp(-448.0, 448.0)
    return t.to(torch.float8_e4m3fn).view(torch.uint8).numpy()


def e4m3_to_f(b):
    return torch.from_numpy(np.ascontiguousarray(b)).view(torch.float8_e4m3fn).float().numpy()


def make_tape(rows, kch):
    m_tiles = (rows + 127) // 128
    m_tiles += m_tiles & 1
    cap = GUARD + m_tiles * 128 + 136
    return dict(rows=rows, m_tiles=m_tiles, cap=cap, kch=kch, kch_stride=cap * 16, part_stride=cap * 16 * kch,
                buf=np.zeros(2 * cap * 16 * kch, np.uint8))


def weight_scale(W):
    m, e = np.frexp(np.float32(np.abs(W).max()))               # weight_scale_kernel
    return np.float32(np.ldexp(1.0, 1 - int(e))), np.float32(np.ldexp(1.0, int(e) - 1))


def pack_b_f16f8(W, n_tiles, stages, BN, kind, K, sw):
    """pack_b_f16f8_kernel, vectorised over idx."""
    out = np.zeros(n_tiles * stages * 8 * BN * 16, np.uint8)
    blk, half = 8 * BN * 16, stages // 2
    idx = np.arange(n_tiles * BN * K, dtype=np.int64)
    k, n = idx % K, idx // K
    nt, nn = n // BN, n % BN
    if kind == 3:
        v = W.reshape(-1)[n * 4736 + (k % 128) * 37 + k // 128]
    else:
        v = W.reshape(-1)[n * K + k]
    v = (v * sw).astype(np.float32)
    h = v.astype(np.float16)
    r = v - h.astype(np.float32)
    s, kr = k // 64, k % 64
    corr = (nt * stages + s) * blk
    mainb = (nt * stages + half + s) * blk
    o16 = mainb + (((kr // 8) * BN + nn) * 8 + kr % 8) * 2
    hb = h.view(np.uint16)
    out[o16] = (hb & 0xFF).astype(np.uint8)
    out[o16 + 1] = (hb >> 8).astype(np.uint8)
    o8 = ((kr // 16) * BN + nn) * 16 + kr % 16
    out[corr + o8] = e4m3(v * 2.0 ** EWH)
    out[corr + blk // 2 + o8] = e4m3(r * 2.0 ** EWL)
    return out


def split16(y, signed=False):
    """split16_f16f8 on [..., 16] -> (fp16 bytes [..., 32], lo8 [..., 16], hi8 [..., 16])."""
    a = np.minimum(y.astype(np.float32), np.float32(65504.0))
    if signed:
        a = np.maximum(a, np.float32(-65504.0))
    h = a.astype(np.float16)
    lo = e4m3((a - h.astype(np.float32)) * 2.0 ** EXL)
    hi = e4m3(a * 2.0 ** EXH)
    return h.view(np.uint8).reshape(*y.shape[:-1], 32), lo.reshape(y.shape), hi.reshape(y.shape)


def write_fc_tape(tape, y, N):
    """tapgemm epilogue, EPI_FC_TAPE with F8: row = window, thread chunks of 32 columns -> 2 x 16."""
    for row in range(y.shape[0]):
        off = (row + GUARD) * 16
        for k16 in range(N // 16):
            f, lo, hi = split16(y[row, 16 * k16: 16 * k16 + 16])
            d16 = (2 * k16) * tape["kch_stride"] + off
            tape["buf"][d16: d16 + 16] = f[:16]
            tape["buf"][d16 + tape["kch_stride"]: d16 + tape["kch_stride"] + 16] = f[16:]
            d8 = tape["part_stride"] + k16 * tape["kch_stride"] + off
            tape["buf"][d8: d8 + 16] = lo
            d8h = d8 + (N // 16) * tape["kch_stride"]
            tape["buf"][d8h: d8h + 16] = hi


def write_block2_tape(tape, y4):
    """block2_kernel<true> epi2: y4[w][to][128 channels] (pooled); thread = (h, c) 32-channel group, hh = 16-channel half."""
    for w in range(y4.shape[0]):
        for to in range(37):
            for h in range(2):
                for c in range(2):
                    for hh in range(2):
                        ch0 = h * 64 + c * 32 + hh * 16
                        f, lo, hi = split16(y4[w, to, ch0: ch0 + 16])
                        d8 = tape["part_stride"] + (to * 8 + h * 4 + c * 2 + hh) * tape["kch_stride"] + (w + GUARD) * 16
                        tape["buf"][d8: d8 + 16] = lo                                  # odd lane
                        d8h = d8 + (37 * 8) * tape["kch_stride"]
                        tape["buf"][d8h: d8h + 16] = hi
                        d16 = (to * 16 + h * 8 + c * 4 + hh * 2) * tape["kch_stride"] + (w + GUARD) * 16
                        tape["buf"][d16: d16 + 16] = f[:16]                             # even lane
                        tape["buf"][d16 + tape["kch_stride"]: d16 + tape["kch_stride"] + 16] = f[16:]


# Issue plans.  By default computed here with the formulas of f8_fc_mma / f8_conv_mma (csrc/dce_tc.cuh); the CPU test
# replaces them with the tables tools/host_check_f16f8.cu dumps from the CUDA source, so that the emulated MMAs below
# execute the kernels' own plan.  A plan entry is (a_off, b_off, e4m3, mode), mode 0 = overwrite, 1 = accumulate,
# 2 = accumulate after scaling D by 2^-15.
PLANS = {}


def fc_plan(s, stages, i, a_part, b_part, b_tapch):
    key = ("fc", stages, b_tapch // 16)
    if key in PLANS:
        return tuple(int(v) for v in PLANS[key][s][i])
    half = stages // 2
    if s < half:
        kk, prod = i >> 1, i & 1
        return (16 + prod * a_part + 2 * kk * SLAB, prod * b_part + 2 * kk * b_tapch, 1, 0 if (s == 0 and i == 0) else 1)
    return (16 + 2 * i * SLAB, 2 * i * b_tapch, 0, 2 if (s == half and i == 0) else 1)


def conv_plan(s, half, tap, i, C, cout):
    key = ("conv", half, C, cout)
    if key in PLANS:
        return tuple(int(v) for v in PLANS[key][s][tap][i])
    if s < half:
        g16 = 2 * s
        a = ((C // 8 + C // 16 + g16) if i else (C // 8 + g16)) * SLAB + tap * 16
        return (a, (i * 3 + tap) * 32 * cout, 1, 0 if (s == 0 and tap == 0 and i == 0) else 1)
    return (2 * (2 * (s - half) + i) * SLAB + tap * 16, (tap * 2 + i) * 32 * cout, 0, 2 if (s == half and tap == 0 and i == 0) else 1)


def apply_mma(D, a, b, mode):
    if mode == 0:
        D[:] = 0.0
    elif mode == 2:
        D *= 2.0 ** -SCALE_D
    D += a @ b.T


def gather(smem, start, lbo, rows, elem_bytes):
    """Operand of one MMA through its SWIZZLE_NONE K-major descriptor: two 16-byte K chunks, SBO = 128 (row pitch 16)."""
    out = []
    for kc in range(2):
        raw = np.stack([smem[start + kc * lbo + r * 16: start + kc * lbo + r * 16 + 16] for r in range(rows)])
        out.append(raw.view(np.float16).astype(np.float64) if elem_bytes == 2 else e4m3_to_f(raw).astype(np.float64))
    return np.concatenate(out, axis=1)                          # [rows][K of the MMA]


def tile(tape, wpk, m, n, stages, BN, MT):
    """One CTA tile of tapgemm_kernel<BN, 1, 4, *, *, MT, 0, F8 = 1>: returns the MT accumulators [MT][128][BN]."""
    KSA = 4
    A_PART, A_TILE, B_TAPCH = KSA * SLAB, 2 * KSA * SLAB, BN * 16
    A_BYTES, B_PART = MT * A_TILE, KSA * B_TAPCH
    B_BYTES = 2 * B_PART
    half = stages // 2
    D = np.zeros((MT, 128, BN))
    a_row = (128 * m + GUARD - 1) * 16
    for s in range(stages):
        st = np.zeros(A_BYTES + B_BYTES, np.uint8)
        for c in range(MT * 2 * KSA):                           # producer: A slabs
            mt, part, j = c // (2 * KSA), (c // KSA) & 1, c % KSA
            if s < half:
                src_off = tape["part_stride"] + (part * 2 * stages + s * KSA + j) * tape["kch_stride"]
            else:
                src_off = ((s - half) * 2 * KSA + part * KSA + j) * tape["kch_stride"]
            src = a_row + mt * 2048 + src_off
            dst = mt * A_TILE + part * A_PART + j * SLAB
            st[dst: dst + SLAB] = tape["buf"][src: src + SLAB]
        wsrc = (n * stages + s) * B_BYTES                       # producer: the B block
        st[A_BYTES: A_BYTES + B_BYTES] = wpk[wsrc: wsrc + B_BYTES]
        for mt in range(MT):                                    # issuers: the plan of f8_fc_mma
            for i in range(4):
                a_off, b_off, e4m3_, mode = fc_plan(s, stages, i, A_PART, B_PART, B_TAPCH)
                eb = 1 if e4m3_ else 2
                a = gather(st, mt * A_TILE + a_off, SLAB, 128, eb)
                b = gather(st, A_BYTES + b_off, B_TAPCH, BN, eb)
                apply_mma(D[mt], a, b, mode)
    return D


def check(rows=5, seed=0):
    """Returns the norm-wise errors of (fc.0 fed by the block2 writer, fc.3 fed by fc.0's writer)."""
    from deep_contact_estimator_b200 import synth
    P = {k: v.numpy() for k, v in synth.make_params(seed).items()}
    rng = np.random.default_rng(seed + 1)
    relu = lambda v: np.maximum(v, 0.0)
    # ---- fc.0: X4 written by block2_kernel<true>, K order k' = t*128 + c (flatten c*37 + t of src/contact_cnn.py:64)
    y4 = relu(rng.standard_normal((rows, 37, 128))).astype(np.float32)          # pooled conv4 output [w][t][c]
    x4 = make_tape(rows, 592)
    write_block2_tape(x4, y4)
    W0 = P["fc.0.weight"]
    sw0, inv0 = weight_scale(W0)
    wp0 = pack_b_f16f8(W0, 8, 148, 256, 3, 4736, sw0)
    want0 = y4.transpose(0, 2, 1).reshape(rows, -1).astype(np.float64) @ W0.astype(np.float64).T      # reference flatten
    errs = []
    h1_full = relu(want0 + P["fc.0.bias"]).astype(np.float32)
    for n in (0, 7):
        D = tile(x4, wp0, 0, n, 148, 256, 2)
        got = D[0][:rows] * float(inv0)
        ref = want0[:, n * 256:(n + 1) * 256]
        errs.append(np.abs(got - ref).max() / np.abs(ref).max())
        assert np.abs(D[1]).max() == 0.0                       # the second M-tile holds only padding rows here
    # ---- fc.3: H1 written by fc.0's epilogue
    h1 = make_tape(rows, 256)
    write_fc_tape(h1, h1_full, 2048)
    W1 = P["fc.3.weight"]
    sw1, inv1 = weight_scale(W1)
    wp1 = pack_b_f16f8(W1, 4, 64, 128, 4, 2048, sw1)
    want1 = h1_full.astype(np.float64) @ W1.astype(np.float64).T
    e3 = []
    for n in range(4):
        D = tile(h1, wp1, 0, n, 64, 128, 1)
        got = D[0][:rows] * float(inv1)
        ref = want1[:, n * 128:(n + 1) * 128]
        e3.append(np.abs(got - ref).max() / np.abs(ref).max())
    return max(errs), max(e3)


# ---------------------------------------------------------------------------------------------------------
# conv_f16f8: block1's X2 writer, pack_conv_f16f8_kernel and block2_kernel<true, true> (slabs, weight blocks, issuer)
# ---------------------------------------------------------------------------------------------------------
def pack_conv_f16f8(W, cout, cin, cin_pad, sw):
    """pack_conv_f16f8_kernel: W[cout][cin][3] -> 2 * cin_pad / 32 blocks of 192 * cout bytes."""
    G, blk = cin_pad // 32, 192 * cout
    out = np.zeros(2 * G * blk, np.uint8)
    idx = np.arange(cout * cin_pad * 3, dtype=np.int64)
    tap, c, n = idx % 3, (idx // 3) % cin_pad, idx // (3 * cin_pad)
    v = np.where(c < cin, W.reshape(-1)[(n * cin + np.minimum(c, cin - 1)) * 3 + tap], 0.0).astype(np.float32) * sw
    h = v.astype(np.float16)
    r = v - h.astype(np.float32)
    g, cr = c // 32, c % 32
    o16 = (G + g) * blk + ((tap * 4 + cr // 8) * cout + n) * 16 + (cr % 8) * 2
    hb = h.view(np.uint16)
    out[o16] = (hb & 0xFF).astype(np.uint8)
    out[o16 + 1] = (hb >> 8).astype(np.uint8)
    o8 = g * blk + ((tap * 2 + cr // 16) * cout + n) * 16 + cr % 16
    out[o8] = e4m3(v * 2.0 ** EWH)
    out[o8 + blk // 2] = e4m3(r * 2.0 ** EWL)
    return out


def write_x2_tape(tape, y2):
    """block1_kernel<.., 1> epi2: y2[w][75][64] pooled conv2 output -> X2 tape (76 rows per window, row 75 = guard)."""
    for w in range(y2.shape[0]):
        for t in range(75):
            orow = w * 76 + t
            for h in range(2):
                for hh in range(2):
                    f, lo, hi = split16(y2[w, t, h * 32 + hh * 16: h * 32 + hh * 16 + 16])
                    d8 = tape["part_stride"] + (h * 2 + hh) * tape["kch_stride"] + (orow + GUARD) * 16
                    tape["buf"][d8: d8 + 16] = lo
                    tape["buf"][d8 + 4 * tape["kch_stride"]: d8 + 4 * tape["kch_stride"] + 16] = hi
                    d16 = (h * 4 + hh * 2) * tape["kch_stride"] + (orow + GUARD) * 16
                    tape["buf"][d16: d16 + 16] = f[:16]
                    tape["buf"][d16 + tape["kch_stride"]: d16 + tape["kch_stride"] + 16] = f[16:]


def conv_blocks(slab, wimg, kch_total):
    """stage_mmas of block2_kernel<true, true> over all blocks of one conv: slab = uint8 [(kch_total * 2) * SLAB]."""
    D = np.zeros((128, 128))
    half, blk = kch_total // 4, 24576
    for s in range(2 * half):
        for tap in range(3):
            for i in range(2):                                  # the plan of f8_conv_mma
                a_off, b_off, e4m3_, mode = conv_plan(s, half, tap, i, kch_total * 8, 128)
                eb = 1 if e4m3_ else 2
                apply_mma(D, gather(slab, a_off, SLAB, 128, eb), gather(wimg, s * blk + b_off, 2048, 128, eb), mode)
    return D


def check_block2(windows=2, seed=0):
    """X2 (block1 writer) -> conv3 -> slabB -> conv4 -> pool, tile by tile; returns the norm-wise error of the pooled output."""
    from deep_contact_estimator_b200 import synth
    import torch.nn.functional as F
    P = synth.make_params(seed)
    rng = np.random.default_rng(seed + 2)
    y2 = np.maximum(rng.standard_normal((windows, 75, 64)), 0.0).astype(np.float32)
    x2 = make_tape(windows * 76, 8)
    write_x2_tape(x2, y2)
    W3, W4 = P["block2.0.weight"].numpy(), P["block2.2.weight"].numpy()
    b3, b4 = P["block2.0.bias"].numpy(), P["block2.2.bias"].numpy()
    (sw3, inv3), (sw4, inv4) = weight_scale(W3), weight_scale(W4)
    w3i, w4i = pack_conv_f16f8(W3, 128, 64, 64, sw3), pack_conv_f16f8(W4, 128, 128, 128, sw4)
    with torch.no_grad():
        t = torch.from_numpy(y2).double().permute(0, 2, 1)
        a3 = F.relu(F.conv1d(t, P["block2.0.weight"].double(), P["block2.0.bias"].double(), padding=1))
        a4 = F.relu(F.conv1d(a3, P["block2.2.weight"].double(), P["block2.2.bias"].double(), padding=1))
        want = F.max_pool1d(a4, 2, 2).permute(0, 2, 1).numpy()                  # [w][37][128]
    NR = windows * 76
    got = np.zeros_like(want)
    seen = np.zeros(want.shape[:2], bool)
    for tile_i in range((NR + 123) // 124):
        b = tile_i * 124
        slabA = np.zeros(16 * SLAB, np.uint8)
        src = (b - 3 + GUARD) * 16                                            # slabA loader: 16 bulk copies of 2080 B
        for c in range(16):
            o = src + (c >> 3) * x2["part_stride"] + (c & 7) * x2["kch_stride"]
            slabA[c * SLAB:(c + 1) * SLAB] = x2["buf"][o: o + SLAB]
        D3 = conv_blocks(slabA, w3i, 8)
        slabB = np.zeros(32 * SLAB, np.uint8)
        for rit in range(128):                                                 # epi1
            r = b - 2 + rit
            valid = r >= 0 and (r % 76) < 75
            y = np.maximum(D3[rit] * float(inv3) + b3, 0.0).astype(np.float32) if valid else np.zeros(128, np.float32)
            for h in range(2):
                for c in range(2):
                    for hh in range(2):
                        ch0 = h * 64 + c * 32 + hh * 16
                        f, lo, hi = split16(y[ch0: ch0 + 16])
                        d16 = (h * 8 + c * 4 + hh * 2) * SLAB + (rit + 1) * 16
                        slabB[d16: d16 + 16] = f[:16]
                        slabB[d16 + SLAB: d16 + SLAB + 16] = f[16:]
                        d8 = (16 + h * 4 + c * 2 + hh) * SLAB + (rit + 1) * 16
                        slabB[d8: d8 + 16] = lo
                        slabB[d8 + 8 * SLAB: d8 + 8 * SLAB + 16] = hi
        D4 = conv_blocks(slabB, w4i, 16)
        y4 = np.maximum(D4 * float(inv4) + b4, 0.0)
        for rit in range(2, 126, 2):                                           # epi2: pool pairs (even, odd) rows
            r = b - 2 + rit
            if r < 0 or r >= NR:
                continue
            w, to = r // 76, (r % 76) >> 1
            if to < 37:
                got[w, to] = np.maximum(y4[rit], y4[rit + 1])
                seen[w, to] = True
    assert seen.all()
    return np.abs(got - want).max() / np.abs(want).max()


def conv64_mmas(slab, wimg):
    """conv_mmas of block1_kernel<.., 3>: 64 -> 64 channels, resident image [e4m3 g0][e4m3 g1][fp16 g0][fp16 g1]."""
    D = np.zeros((128, 64))
    for s in range(4):
        for tap in range(3):
            for i in range(2):                                  # the plan of f8_conv_mma with half = 2
                a_off, b_off, e4m3_, mode = conv_plan(s, 2, tap, i, 64, 64)
                eb = 1 if e4m3_ else 2
                apply_mma(D, gather(slab, a_off, SLAB, 128, eb), gather(wimg, s * 12288 + b_off, 1024, 64, eb), mode)
    return D


def put_row64(slab, row, y64):
    """One 64-channel row in the slab format of block1: fp16 chunks 0..7, lo8 8..11, hi8 12..15."""
    for c16 in range(4):
        f, lo, hi = split16(y64[16 * c16: 16 * c16 + 16], signed=True)
        slab[(2 * c16) * SLAB + row * 16: (2 * c16) * SLAB + row * 16 + 16] = f[:16]
        slab[(2 * c16 + 1) * SLAB + row * 16: (2 * c16 + 1) * SLAB + row * 16 + 16] = f[16:]
        slab[(8 + c16) * SLAB + row * 16: (8 + c16) * SLAB + row * 16 + 16] = lo
        slab[(12 + c16) * SLAB + row * 16: (12 + c16) * SLAB + row * 16 + 16] = hi


def check_block1(windows=1, seed=0):
    """windows -> converter -> conv1 -> slab1 -> conv2 -> pool -> X2 tape (block1_kernel<false, 3>), decoded and compared."""
    from deep_contact_estimator_b200 import synth
    import torch.nn.functional as F
    P = synth.make_params(seed)
    x = synth.make_windows(windows, seed=seed + 3)                              # [w][150][54], z-scored, signed
    W1, W2 = P["block1.0.weight"].numpy(), P["block1.2.weight"].numpy()
    b1, b2 = P["block1.0.bias"].numpy(), P["block1.2.bias"].numpy()
    (sw1, inv1), (sw2, inv2) = weight_scale(W1), weight_scale(W2)
    w1i, w2i = pack_conv_f16f8(W1, 64, 54, 64, sw1), pack_conv_f16f8(W2, 64, 64, 64, sw2)
    with torch.no_grad():
        t = x.double().permute(0, 2, 1)
        a1 = F.relu(F.conv1d(t, P["block1.0.weight"].double(), P["block1.0.bias"].double(), padding=1))
        a2 = F.relu(F.conv1d(a1, P["block1.2.weight"].double(), P["block1.2.bias"].double(), padding=1))
        want = F.max_pool1d(a2, 2, 2).permute(0, 2, 1).numpy()                  # [w][75][64]
    xn = x.numpy()
    NR = windows * 152
    x2 = make_tape(windows * 76, 8)
    cap_rows = x2["m_tiles"] * 128
    for tile_i in range((NR + 123) // 124):
        b = tile_i * 124
        slab0 = np.zeros(16 * SLAB, np.uint8)
        for srow in range(130):                                                # converters (rows 128, 129 by 8 threads)
            r = b - 3 + srow
            y = np.zeros(64, np.float32)
            if 0 <= r < NR and r % 152 < 150:
                y[:54] = xn[r // 152, r % 152]
            put_row64(slab0, srow, y)
        D1 = conv64_mmas(slab0, w1i)
        slab1 = np.zeros(16 * SLAB, np.uint8)
        for rit in range(128):                                                 # epi1 -> slab1 row rit + 1
            r = b - 2 + rit
            valid = r >= 0 and r % 152 < 150
            y = np.maximum(D1[rit] * float(inv1) + b1, 0.0).astype(np.float32) if valid else np.zeros(64, np.float32)
            put_row64(slab1, rit + 1, y)
        D2 = conv64_mmas(slab1, w2i)
        y2 = np.maximum(D2 * float(inv2) + b2, 0.0).astype(np.float32)
        for rit in range(0, 128, 2):                                           # epi2: pooled pair (rit, rit + 1)
            r = b - 2 + rit
            orow = r >> 1
            valid = r >= 0 and ((r % 152) >> 1) < 75
            store = ((2 <= rit < 126) or (tile_i == 0 and rit < 2)) and orow < cap_rows
            if not store:
                continue
            m = np.maximum(y2[rit], y2[rit + 1]) if valid else np.zeros(64, np.float32)
            for h in range(2):
                for hh in range(2):
                    f, lo, hi = split16(m[h * 32 + hh * 16: h * 32 + hh * 16 + 16])
                    d8 = x2["part_stride"] + (h * 2 + hh) * x2["kch_stride"] + (orow + GUARD) * 16
                    x2["buf"][d8: d8 + 16] = lo
                    x2["buf"][d8 + 4 * x2["kch_stride"]: d8 + 4 * x2["kch_stride"] + 16] = hi
                    d16 = (h * 4 + hh * 2) * x2["kch_stride"] + (orow + GUARD) * 16
                    x2["buf"][d16: d16 + 16] = f[:16]
                    x2["buf"][d16 + x2["kch_stride"]: d16 + x2["kch_stride"] + 16] = f[16:]
    # decode the X2 tape: value = fp16 + lo8 * 2^-12; guard row 75 of every window must be zero
    cap, ks, ps = x2["cap"], x2["kch_stride"], x2["part_stride"]
    f16 = x2["buf"][:ps].view(np.float16).reshape(8, cap, 8).astype(np.float64)
    lo = e4m3_to_f(x2["buf"][ps: ps + 4 * ks]).reshape(4, cap, 16).astype(np.float64) * 2.0 ** -EXL
    val = f16.transpose(1, 0, 2).reshape(cap, 64) + lo.transpose(1, 0, 2).reshape(cap, 64)
    got = val[GUARD: GUARD + windows * 76].reshape(windows, 76, 64)
    assert np.abs(got[:, 75]).max() == 0.0 and np.abs(val[GUARD - 1]).max() == 0.0
    return np.abs(got[:, :75] - want).max() / np.abs(want).max()


if __name__ == "__main__":
    print(f"block1 (converter -> conv1 -> slab1 -> conv2 -> pool -> X2): {check_block1():.2e}")
    print(f"block2 (block1 X2 writer -> conv3 -> slabB -> conv4 -> pool): {check_block2():.2e}")
    e0, e3 = check()
    print(f"fc.0 (block2 writer -> F8 tile): {e0:.2e}   fc.3 (fc.0 writer -> F8 tile): {e3:.2e}")
