"""CPU model of the warp-collective transposed fp32 store of the training GEMM epilogues
(csrc/ptx.cuh::warp_store_rows_f32_v4): the lane -> (row, column) maps of the way into the shared-memory tile and of
the way out cover every element of a 32 x 32 chunk exactly once, leave as 64 contiguous bytes per row, and touch
disjoint shared-memory banks per quarter warp (the reason for the 80-byte row stride and the (r, r + 4) row pairing).
The constants are read from the header, so a change of the tile geometry has to keep these properties."""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PTX = os.path.join(HERE, "..", "robot_aware_control_b200", "csrc", "ptx.cuh")


def _const(name):
    m = re.search(rf"constexpr int {name} = ([^;]+);", open(PTX).read())
    assert m, name
    return m.group(1)


def _geometry():
    row_floats = int(_const("kStageRowFloats"))
    warp_bytes = eval(_const("kStageWarpBytes"), {"kStageRowFloats": row_floats})
    return row_floats, warp_bytes


def _lane_row_col(lane):
    """the way out: rsel / c4 of warp_store_rows_f32_v4"""
    return (lane >> 3) + 4 * ((lane >> 2) & 1), (lane & 3) * 4


def test_tile_geometry_from_header():
    row_floats, warp_bytes = _geometry()
    assert row_floats >= 18, "16 data columns + one 64-bit row offset"
    assert (row_floats * 4) % 16 == 0, "rows stay 16-byte aligned for the 128-bit accesses"
    assert warp_bytes == 32 * row_floats * 4
    src = open(PTX).read()
    assert "const int rsel = (lane >> 3) + 4 * ((lane >> 2) & 1), c4 = (lane & 3) * 4;" in src, \
        "warp_store_rows_f32_v4 changed its lane map: update _lane_row_col"


def test_every_element_stored_once_in_64_byte_segments():
    row_floats, _ = _geometry()
    ncols = 1000  # row stride of the destination in floats
    acc = np.arange(32 * 32, dtype=np.float32).reshape(32, 32)  # acc[lane][column]: TMEM gives a lane one row
    out = np.full((32, ncols), -1.0, np.float32)
    writes = np.zeros((32, ncols), np.int32)
    for h in range(2):
        stage = np.zeros((32, row_floats), np.float32)
        stage[:, :16] = acc[:, 16 * h:16 * h + 16]          # the way in: lane == row, four 128-bit stores
        for i in range(4):                                  # the way out: 4 store instructions per half
            segs = {}
            for lane in range(32):
                rsel, c4 = _lane_row_col(lane)
                row = 8 * i + rsel
                col = 16 * h + c4
                out[row, col:col + 4] = stage[row, c4:c4 + 4]
                writes[row, col:col + 4] += 1
                segs.setdefault(row, []).append(col)
            assert len(segs) == 8, "one instruction serves eight rows"
            for cols in segs.values():
                assert sorted(cols) == [16 * h, 16 * h + 4, 16 * h + 8, 16 * h + 12], "64 contiguous bytes per row"
    assert np.array_equal(out[:, :32], acc)
    assert np.all(writes[:, :32] == 1) and np.all(writes[:, 32:] == 0)


def test_shared_memory_accesses_are_conflict_free_per_quarter_warp():
    row_floats, _ = _geometry()
    row_bytes = row_floats * 4

    def groups(addresses):  # 16-byte bank groups (8 of them = the 32 banks) of a quarter warp's 128-bit accesses
        return [(a // 16) % 8 for a in addresses]

    for q in range(4):  # the way in: lane L writes 16 bytes at row L, column group j (same j for all lanes)
        lanes = range(8 * q, 8 * q + 8)
        for j in range(4):
            g = groups([lane * row_bytes + 16 * j for lane in lanes])
            assert len(set(g)) == 8, ("store", q, j, g)
    for q in range(4):  # the way out: four lanes per row, rows r and r + 4 in one quarter warp
        lanes = range(8 * q, 8 * q + 8)
        for i in range(4):
            addr = []
            for lane in lanes:
                rsel, c4 = _lane_row_col(lane)
                addr.append((8 * i + rsel) * row_bytes + 4 * c4)
            g = groups(addr)
            assert len(set(g)) == 8, ("load", q, i, g)
