"""CPU checks of the index algebra the fused kernels rely on (constants are read from the CUDA sources, so an edit
that breaks an invariant fails here before it reaches a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ofdm-course_b200", "csrc")


def _define(src, name):
    m = re.search(r"#define\s+%s\s+(\d+)" % name, src)
    assert m, name
    return int(m.group(1))


def _read(name):
    return open(os.path.join(CSRC, name)).read()


def test_rx4096_exchange_layout_and_pass_c_reads_are_conflict_free():
    src = _read("chain_rx4096.cu")
    XROW, XGRP, XBUF = _define(src, "XROW"), _define(src, "XGRP"), _define(src, "XBUF")
    assert XBUF == 16 * XROW and XROW >= 15 * XGRP + 16 and XGRP >= 16 and XGRP % 2 == 0      # 16-byte aligned groups of 16 samples
    # pass C: lane (k1c, k2c) reads eight float4 at samples k1c*XROW + k2c*XGRP + 2j; a 128-bit request is served per
    # quarter warp, whose eight lanes hold eight consecutive k2 of one k1: their 16-byte bank groups must differ
    for tid0 in range(0, 256, 8):
        lanes = range(tid0, tid0 + 8)
        for j in range(8):
            groups = set()
            for t in lanes:
                k1c = (t >> 6) + 4 * ((t >> 3) & 3)
                k2c = ((t >> 5) & 1) * 8 + (t & 7)
                byte = 8 * (k1c * XROW + k2c * XGRP + 2 * j)
                assert byte % 16 == 0
                groups.add((byte // 16) % 8)
            assert len(groups) == 8


def test_rx4096_inverse_estimate_slots_match_the_pass_c_owner():
    # Hinv is stored so that thread tid finds 1/H of its carriers kq + 256c at Hinv[tid + 256c] (chain_rx4096.cu, `keep`)
    owner = {}
    for tid in range(256):
        k1c = (tid >> 6) + 4 * ((tid >> 3) & 3)
        k2c = ((tid >> 5) & 1) * 8 + (tid & 7)
        for c in range(4):
            owner[k1c + 16 * k2c + 256 * c] = tid + 256 * c
    assert sorted(owner) == list(range(1024))
    for q in range(1024):
        k1, k2 = q & 15, (q >> 4) & 15
        slot = (((k1 & 3) << 6) | ((k2 >> 3) << 5) | ((k1 >> 2) << 3) | (k2 & 7)) + (q & ~255)
        assert slot == owner[q]
    # comb-4 pilots (carriers 0, 4, 8, ... 0-based) fall into whole warps: the ones the kernel lets skip pass C
    pilot_warps = {tid >> 5 for tid in range(256) for c in range(4)
                   if (((tid >> 6) + 4 * ((tid >> 3) & 3)) + 16 * (((tid >> 5) & 1) * 8 + (tid & 7)) + 256 * c) % 4 == 0}
    assert pilot_warps == {0, 1}


def test_tx_scrambler_zero_pad_covers_every_shifted_read():
    src = _read("chain.cu")
    m = re.search(r"#define\s+TXF_PAD\(fw\)\s+\(\(fw\) \+ \(\(fw\) >> (\d+)\) \+ (\d+)\)", src)
    assert m
    sh, add = int(m.group(1)), int(m.group(2))
    for fw in list(range(1, 200)) + [672, 1024, 2048, 3333]:
        pad = fw + (fw >> sh) + add
        for frame_bits in {32 * (fw - 1) + 1, 32 * fw - 7, 32 * fw} - {0}:
            if frame_bits <= 0:
                continue
            s13, s14 = 13, 14
            while s13 < frame_bits:                                   # the doubling loop of tx4096_kernel
                reach = (s14 >> 5) + 1                                # lowest word read: w - q14 - 1 with w = 0
                assert reach <= pad, (fw, frame_bits, s14)
                s13, s14 = 2 * s13, 2 * s14


def test_channel_history_regeneration_fills_exactly_the_history_slots():
    src = _read("channel.cu")
    tile, chunk = _define(src, "CH_TILE"), _define(src, "CH_CHUNK")
    assert tile % 512 == 0 and chunk >= 1 and _define(src, "CH_MAXD") - 1 <= tile
    c0 = 4 * tile
    for D in range(1, 60):
        slots = {}
        q = 0
        while 2 * q < D - 1:                                         # pair c0/2 - 1 - q covers samples c0 - 2q - 2, c0 - 2q - 1
            pr = (c0 >> 1) - 1 - q
            j1 = D - 2 - 2 * q
            slots[j1] = 2 * pr + 1
            if j1 >= 1:
                slots[j1 - 1] = 2 * pr
            q += 1
        assert sorted(slots) == list(range(D - 1))
        assert all(slots[j] == c0 - (D - 1) + j for j in slots)      # slot j holds sample c0 - (D-1) + j
