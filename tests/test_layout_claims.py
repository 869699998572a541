"""Shared-memory bank arithmetic behind two layout decisions of the kernels, restated in Python so that
the claims are checked where they are made: the swizzled staging rows of k_kron_mma (kernels_camera.cu) and the
interleaved landmark record of the camera-major passes (povar_internal.h, kLmRec)."""
import itertools


def swz(row):
    return 2 * ((row >> 1) & 1) + ((row >> 2) & 1)


def store_groups(swizzled):
    """16-byte bank groups (8 per 128 bytes) hit by the eight lanes of a quarter warp that store chunk q of their own
    64-byte row (STS.128)."""
    worst = 0
    for quarter, q in itertools.product(range(4), range(4)):
        groups = {}
        for lane in range(8 * quarter, 8 * quarter + 8):
            pos = q ^ swz(lane) if swizzled else q
            g = (4 * lane + pos) % 8
            groups[g] = groups.get(g, 0) + 1
        worst = max(worst, max(groups.values()))
    return worst


def fragment_read_units(swizzled):
    """8-byte units (16 per 128 bytes) hit by a half warp that reads column fr of row 4 j + fk (LDS.64), lane = 4 fr + fk."""
    worst = 0
    for j, half in itertools.product(range(8), range(2)):
        units = {}
        for lane in range(16 * half, 16 * half + 16):
            fr, fk = lane >> 2, lane & 3
            row = 4 * j + fk
            chunk = (fr >> 1) ^ swz(row) if swizzled else fr >> 1
            u = (8 * row + 2 * chunk + (fr & 1)) % 16
            units[u] = units.get(u, 0) + 1
        worst = max(worst, max(units.values()))
    return worst


def test_kron_staging_swizzle_is_conflict_free():
    assert store_groups(swizzled=False) == 4 and fragment_read_units(swizzled=False) == 2   # what bounded the kernel
    assert store_groups(swizzled=True) == 1 and fragment_read_units(swizzled=True) == 1


def test_kernel_uses_the_same_swizzle_as_the_model():
    # the kernel computes swz of the row it writes from the lane, and of the row 4 j + fk it reads from (fk, j)
    for lane in range(32):
        assert swz(lane) == 2 * ((lane >> 1) & 1) + ((lane >> 2) & 1)
    for j, fk in itertools.product(range(8), range(4)):
        assert swz(4 * j + fk) == 2 * (fk >> 1) + (j & 1)


def test_landmark_record_halves_are_whole_sectors():
    # [X0 X1 H0 H1 | X2 X3 H2 H3]: lane j of an entry needs X[2j], X[2j+1], H[2j], H[2j+1] -- one aligned 32-byte piece
    x_at, h_at = {0: 0, 1: 1, 2: 4, 3: 5}, {0: 2, 1: 3, 2: 6, 3: 7}
    for j in range(2):
        need = sorted([x_at[2 * j], x_at[2 * j + 1], h_at[2 * j], h_at[2 * j + 1]])
        assert need == list(range(4 * j, 4 * j + 4))
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "povar_b200", "csrc", "povar_internal.h")).read()
    m = re.search(r"kLmRecX0 = (\d+), kLmRecH0 = (\d+), kLmRecX2 = (\d+), kLmRecH2 = (\d+)", src)
    assert m and [int(v) for v in m.groups()] == [x_at[0], h_at[0], x_at[2], h_at[2]]
