"""Shared-memory operand layouts the tcgen05 kernels rely on, pinned to what the hardware was MEASURED to read
(tests/golden/umma_probe_b200.txt = profiles/r2_probe_umma.txt, produced on a B200 by scripts/probe_umma_mn.py: a one-MMA probe whose result is the
shared-memory float index read as B(n, k)).  CPU only: the probe output is a committed fixture."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(HERE, "golden", "umma_probe_b200.txt")  # a copy of profiles/r2_probe_umma.txt


def _probe_rows(header_prefix):
    lines = open(PROBE).read().split("\n")
    i = next(j for j, l in enumerate(lines) if l.startswith(header_prefix))
    return [[int(v) for v in lines[i + 1 + k].split(":")[1].split()] for k in range(8)]


def mn_major_atom32_index(n, k, lbo_bytes, sbo_bytes):
    """Float index of B(n, k) for an MN-major kind::tf32 operand in the 128-byte swizzle with 32-byte atoms (UMMA layout
    type 1), as csrc/wgrad_tc.cu assumes it: 32 MN elements per 128-byte row, 32-byte units XOR (k & 3), 4 k rows per
    group (stride SBO), the next 32 MN elements at LBO."""
    return (n // 32) * (lbo_bytes // 4) + (k // 4) * (sbo_bytes // 4) + (k % 4) * 32 + ((((n % 32) >> 3) ^ (k & 3)) * 8) + n % 8


def test_mn_major_layout_type_1_matches_the_probe():
    for lbo, sbo in ((4096, 1024), (1024, 4096)):
        rows = _probe_rows(f"MN-major layout=1 lbo={lbo} sbo={sbo}")
        assert len(rows[0]) == 64
        assert rows == [[mn_major_atom32_index(n, k, lbo, sbo) for n in range(64)] for k in range(8)]


def test_other_layouts_read_nothing_for_mn_major_tf32():
    # no-swizzle and the 16-byte-atom swizzles: the tensor core answers with zeros (index 0 everywhere)
    for layout in (0, 2, 4, 6):
        lines = open(PROBE).read().split("\n")
        i = next(j for j, l in enumerate(lines) if l.startswith(f"MN-major layout={layout} "))
        for k in range(8):
            assert set(lines[i + 1 + k].split(":")[1].split()) == {"0"}


def test_wgrad_tc_thread_to_feature_mapping_follows_the_layout():
    """wgrad_tc.cu's rounding warps: thread rt owns the 16-byte chunk at position rt % 8 of row rt / 8 of every 32-row x
    32-feature box (rows 128 bytes apart, contiguous: SBO = 512) and treats it as features `feat .. feat + 3`."""
    for rt in range(256):
        row, cpos = rt >> 3, rt & 7
        feat = ((((rt & 7) >> 1) ^ ((rt >> 3) & 3)) << 3) | ((rt & 1) << 2)  # the kernel's expression
        for e in range(4):
            phys = row * 32 + cpos * 4 + e  # float index inside the box
            assert mn_major_atom32_index(feat + e, row, 4096, 512) == phys


def test_tma_box_swizzle_128b_chunk_model():
    """K-major SWIZZLE_128B boxes (linear.cu, rows_gemm_tc.cu): the 16-byte chunk j of row r sits at position j ^ (r & 7);
    rows_gemm_tc.cu's fix-up thread ft (chunk position ft % 8 of rows ft / 8 + 32 i) therefore always meets logical chunk
    (ft % 8) ^ ((ft / 8) & 7), whatever i."""
    for ft in range(256):
        cpos, r0 = ft & 7, ft >> 3
        kk = (cpos ^ (r0 & 7)) << 2  # the kernel's expression: first of the thread's four K columns
        for i in range(4):
            r = r0 + 32 * i
            logical = cpos ^ (r & 7)
            assert logical * 4 == kk
