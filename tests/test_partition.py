"""Host-side logic of the N-GPU path on CPU: box geometry and the IPC-handle routing between ranks
(world_size 2 and 4 over the gloo backend).  The device side is covered by tests/mgpu_parity.py."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from nextsimdg_b200.partition import BOTTOM, LEFT, OPPOSITE, RIGHT, TOP, Partition, grid_shape


@pytest.mark.parametrize("n,shape", [(1, (1, 1)), (2, (2, 1)), (4, (2, 2)), (8, (4, 2))])
def test_grid_shape(n, shape):
    assert grid_shape(n) == shape


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_boxes_tile_the_domain_and_neighbours_are_symmetric(nranks):
    gnx, gny = 64, 32
    parts = [Partition.strong(r, nranks, gnx, gny) for r in range(nranks)]
    cover = np.zeros((gny, gnx), dtype=int)
    for p in parts:
        cover[p.owned_window()] += 1
        for s in range(4):
            q = p.neighbour[s]
            if q >= 0:
                assert parts[q].neighbour[OPPOSITE[s]] == p.rank
                assert p.ring[s] == 1
            else:  # a side without neighbour lies on the global edge
                assert {BOTTOM: p.y0 == 0, TOP: p.y0 + p.ny == gny, LEFT: p.x0 == 0, RIGHT: p.x0 + p.nx == gnx}[s]
        # the local window is the owned box plus the ring, clipped by nothing (rings exist only towards neighbours)
        ly, lx = p.local_window()
        assert (lx.start, lx.stop) == (p.x0 - p.ring[LEFT], p.x0 + p.nx + p.ring[RIGHT])
        assert (ly.start, ly.stop) == (p.y0 - p.ring[BOTTOM], p.y0 + p.ny + p.ring[TOP])
    assert (cover == 1).all()


def test_crop_state_and_owned_slices():
    gnx, gny = 12, 8
    ms = {"coords": np.arange((gny + 1) * (gnx + 1) * 2, dtype=float).reshape(gny + 1, gnx + 1, 2),
          "hice": np.arange(gny * gnx, dtype=float).reshape(gny, gnx)}
    for r in range(4):
        p = Partition.strong(r, 4, gnx, gny)
        loc = p.crop_state(ms)
        assert loc["hice"].shape == (p.lny, p.lnx) and loc["coords"].shape == (p.lny + 1, p.lnx + 1, 2)
        assert np.array_equal(loc["hice"][p.owned_in_local()], ms["hice"][p.owned_window()])


class _FakeLib:
    """stands in for libnsdg_cuda: records what connect_halos routes where"""

    def __init__(self, rank):
        self.rank, self.connected, self.ready = rank, {}, False

    def nsdg_halo_export(self, h, buf):
        import ctypes

        arr = (ctypes.c_ubyte * 64).from_address(buf.value)
        for i in range(64):
            arr[i] = (self.rank * 7 + i) % 251
        return 0

    def nsdg_halo_connect(self, h, side, buf):
        import ctypes

        self.connected[side] = bytes((ctypes.c_ubyte * 64).from_address(buf.value))
        return 0

    def nsdg_halo_ready(self, h):
        self.ready = True
        return 0


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from nextsimdg_b200.partition import connect_halos, torch_all_gather

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = Partition.strong(rank, world, 32, 32)

    class Dyn:
        _lib = _FakeLib(rank)
        _h = None

    connect_halos(Dyn, part, torch_all_gather(dist))
    expect = {s: bytes(((part.neighbour[s] * 7 + i) % 251) for i in range(64)) for s in range(4) if part.neighbour[s] >= 0}
    q.put((rank, Dyn._lib.connected == expect and Dyn._lib.ready))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ipc_handle_routing_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world * 7 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]


# ---- the reference's partition metadata files ----
REF = os.environ.get("NSDG_REFERENCE_ROOT", "/root/reference")
# run/partition.cdl and core/test/partition_metadata_{2,3}.cdl, restated so that the test also runs where the reference tree is absent
CDL_1 = """netcdf partition { dimensions: P = 1 ; L = 1 ; NX = 30 ; NY = 30 ;
group: bounding_boxes { variables: int domain_x(P) ; data:
 domain_x = 0 ; domain_y = 0 ; domain_extent_x = 30 ; domain_extent_y = 30 ; } }"""


def _ref_text(rel, fallback=None):
    path = os.path.join(REF, rel)
    if os.path.exists(path):
        return open(path).read()
    if fallback is None:
        pytest.skip("reference tree not present")
    return fallback


def test_reads_the_reference_single_box_partition_file():
    from nextsimdg_b200.partition import PartitionFile

    pf = PartitionFile.from_cdl(_ref_text("run/partition.cdl", CDL_1))
    assert (pf.global_nx, pf.global_ny, pf.boxes) == (30, 30, [(0, 0, 30, 30)])
    p = pf.partition(0)
    assert (p.nx, p.ny, p.x0, p.y0, p.neighbour, p.lnx, p.lny) == (30, 30, 0, 0, [-1, -1, -1, -1], 30, 30)


def test_reads_the_reference_two_box_partition_file():
    """core/test/partition_metadata_2.cdl: boxes at domain_x = 0 and 4 of a 10 x 9 grid; the file's bottom / top are +x / -x."""
    from nextsimdg_b200.partition import PartitionFile

    pf = PartitionFile.from_cdl(_ref_text("core/test/partition_metadata_2.cdl"))
    assert (pf.global_nx, pf.global_ny) == (10, 9) and pf.boxes == [(0, 0, 4, 9), (4, 0, 6, 9)]
    assert pf.connectivity["bottom"] == [[(1, 9)], []] and pf.connectivity["top"] == [[], [(0, 9)]]
    pf.check_connectivity()
    p0, p1 = pf.partition(0), pf.partition(1)
    assert p0.neighbour == [-1, 1, -1, -1] and p1.neighbour == [-1, -1, -1, 0]
    assert (p0.lnx, p0.lny, p1.lx0, p1.lnx) == (5, 9, 3, 7)
    cfg_fields = {}

    class Cfg:
        neighbour = [0, 0, 0, 0]

    c = Cfg()
    p1.fill_config(c)
    assert (c.global_nx, c.global_ny, c.box_x0, c.box_y0, c.rank, c.nranks, c.neighbour) == (10, 9, 4, 0, 1, 2, [-1, -1, -1, 0])
    del cfg_fields


def test_reads_the_reference_three_box_partition_file():
    """core/test/partition_metadata_3.cdl:28-66: box 2 faces TWO boxes across one side -- parsed and cross-checked exactly,
    but refused as a device partition (one neighbour per side)."""
    from nextsimdg_b200.partition import LEFT, PartitionFile

    pf = PartitionFile.from_cdl(_ref_text("core/test/partition_metadata_3.cdl"))
    assert (pf.global_nx, pf.global_ny) == (5, 7)
    assert pf.boxes == [(0, 0, 3, 3), (0, 3, 3, 4), (3, 0, 2, 7)]
    assert pf.connectivity["top"][2] == [(0, 3), (1, 4)] and pf.connectivity["right"][0] == [(1, 3)]
    pf.check_connectivity()
    assert sorted(pf.neighbours_geometric(2)[LEFT]) == [(0, 3), (1, 4)]
    with pytest.raises(ValueError):
        pf.partition(2)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_partition_file_round_trip_of_the_regular_box_grids(nranks):
    from nextsimdg_b200.partition import PartitionFile, write_partition_cdl

    parts = [Partition.strong(r, nranks, 64, 32) for r in range(nranks)]
    pf = PartitionFile.from_cdl(write_partition_cdl(parts))
    pf.check_connectivity()
    for r, want in enumerate(parts):
        got = pf.partition(r)
        for k in ("nx", "ny", "x0", "y0", "neighbour", "ring", "lx0", "ly0", "lnx", "lny", "global_nx", "global_ny", "nranks"):
            assert getattr(got, k) == getattr(want, k), (r, k)
