"""Host-side multi-GPU logic on CPU: partition math and a world_size-2 gloo run of the collectives the N>1 path uses
(SURVEY.md §8e).  The GPU side is tests/test_sharding_gpu.py (needs 2 GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_video_shard_partitions_exactly():
    from posetraj_b200.sharding import video_shard
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in video_shard(n, r, world)]
            assert got == list(range(n))
            sizes = [len(video_shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        video_shard(4, 2, 2)


def test_frame_shards_ragged():
    from posetraj_b200.sharding import frame_shards
    assert frame_shards(25, 8) == [(0, 4), (4, 3), (7, 3), (10, 3), (13, 3), (16, 3), (19, 3), (22, 3)]
    assert frame_shards(14, 2) == [(0, 7), (7, 7)]
    assert sum(c for _, c in frame_shards(25, 4)) == 25


def test_temporal_context_rotation_matches_reference_indexing():
    """hidden row (b, s) <- context of batch (b*HW + s) mod B (models/modified_svd.py:152-159 vs :64-66)."""
    from posetraj_b200.sharding import temporal_context_rotation
    B = 2
    for hw in (2880, 720, 180, 45):
        for row in range(B):
            rot = temporal_context_rotation(row, hw, B)
            for s in (0, 1, 2, 43, 44):
                assert (0 * hw + s + rot) % B == (row * hw + s) % B     # local row 0 + rotation == global indexing
    assert temporal_context_rotation(1, 45) == 1 and temporal_context_rotation(1, 2880) == 0


def test_cfg_split_ranks():
    from posetraj_b200.sharding import cfg_split_ranks
    assert cfg_split_ranks(8) == [(0, 1), (2, 3), (4, 5), (6, 7)]
    with pytest.raises(ValueError):
        cfg_split_ranks(3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posetraj_b200.sharding import gather_latents, max_over_ranks, video_shard
    vids = list(video_shard(5, rank, world))
    # pretend-denoise: each rank's "latents" encode the video ids it owns (pad to equal length for the gather)
    lat = torch.full((3, 2, 4, 2, 2), -1.0)
    for i, v in enumerate(vids):
        lat[i] = float(v)
    allv = gather_latents(lat)
    ids = sorted(int(x) for x in allv[:, 0, 0, 0, 0].tolist() if x >= 0)
    t = max_over_ranks(10.0 + rank)
    # the CFG exchange: all_gather_into_tensor of one row per rank reproduces the [2, ...] prediction
    pred = torch.full((6, 4), float(rank))
    both = torch.empty(12, 4)
    dist.all_gather_into_tensor(both, pred)
    q.put((rank, ids, t, both[:6].mean().item(), both[6:].mean().item()))
    dist.destroy_process_group()


def test_gloo_world2_collectives():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=400) for _ in range(world))
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    for rank, ids, t, m0, m1 in res:
        assert ids == [0, 1, 2, 3, 4]
        assert t == 11.0                     # max over ranks
        assert (m0, m1) == (0.0, 1.0)        # rank r's row lands in slot r


def _simulate_exchange(B, F, HW, world, Cc=3):
    """numpy model of the frame<->pixel re-sharding: F-layout on every rank -> pack -> all-to-all (split sizes as in
    ShardedNetPlan) -> unpack -> P-layout; then back.  Checks every element lands where the layouts say it should."""
    import numpy as np
    from posetraj_b200.frame_sharding import f2p_tables, f2p_unpack_tables, p2f_pack_tables, p2f_tables, pixel_shards
    from posetraj_b200.sharding import frame_shards
    fsh, psh = frame_shards(F, world), pixel_shards(HW, world)
    full = np.arange(B * F * HW * Cc, dtype=np.int64).reshape(B, F, HW, Cc)

    def copy(src_t, tables, n_rows):
        out = np.full((n_rows, Cc), -1, dtype=np.int64)
        for s_, d_, n_ in zip(*tables):
            out[d_:d_ + n_] = src_t[s_:s_ + n_]
        return out

    def all_to_all(send_bufs, send_rows):      # send_rows[r][d]; chunks arrive ordered by source rank
        outs = []
        for r in range(world):
            parts = []
            for s_rank in range(world):
                off = sum(send_rows[s_rank][:r])
                parts.append(send_bufs[s_rank][off:off + send_rows[s_rank][r]])
            outs.append(np.concatenate(parts, 0))
        return outs

    f_lay = [full[:, f0:f0 + nf].reshape(B * nf * HW, Cc) for f0, nf in fsh]
    # ---- F -> P
    send = [copy(f_lay[r], f2p_tables(B, nf, HW, psh), B * nf * HW) for r, (f0, nf) in enumerate(fsh)]
    recv = all_to_all(send, [[B * nf * nq for _, nq in psh] for _, nf in fsh])
    p_lay = []
    for r, (p0, npx) in enumerate(psh):
        out = copy(recv[r], f2p_unpack_tables(B, F, npx, fsh), B * F * npx)
        assert np.array_equal(out, full[:, :, p0:p0 + npx].reshape(B * F * npx, Cc)), ("F->P", r)
        p_lay.append(out)
    # ---- P -> F
    send = [copy(p_lay[r], p2f_pack_tables(B, F, npx, fsh), B * F * npx) for r, (p0, npx) in enumerate(psh)]
    recv = all_to_all(send, [[B * cnt * npx for _, cnt in fsh] for _, npx in psh])
    for r, (f0, nf) in enumerate(fsh):
        out = copy(recv[r], p2f_tables(B, nf, HW, psh), B * nf * HW)
        assert np.array_equal(out, f_lay[r]), ("P->F", r)


def test_frame_pixel_exchange_tables():
    for B, F, HW, world in [(2, 14, 45, 2), (2, 25, 144, 8), (2, 5, 6, 3), (1, 3, 7, 3), (2, 25, 9216, 4)]:
        _simulate_exchange(B, F, HW, world)


def _a2a_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posetraj_b200.frame_sharding import AllReduceSum, AllToAllRows
    send_rows = [2, 3] if rank == 0 else [1, 4]          # ragged: rank0 -> (2 rows to r0, 3 to r1); rank1 -> (1, 4)
    recv_rows = [2, 1] if rank == 0 else [3, 4]
    src = torch.arange(sum(send_rows) * 2, dtype=torch.float32).reshape(-1, 2) + 100 * rank
    dst = torch.zeros(sum(recv_rows), 2)
    AllToAllRows(src, dst, send_rows, recv_rows, None).launch(0)
    s = torch.full((4,), float(rank + 1), dtype=torch.float64)
    AllReduceSum(s, None).launch(0)
    q.put((rank, dst.tolist(), s.tolist()))
    dist.destroy_process_group()


def test_gloo_world2_ragged_all_to_all_rows():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_a2a_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (d, s)) for r, d, s in (q.get(timeout=400) for _ in range(2)))
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    # rank 0 receives its own first 2 rows and rank 1's first row; rank 1 receives rank 0's last 3 and its own last 4
    assert res[0][0] == [[0.0, 1.0], [2.0, 3.0], [100.0, 101.0]]
    assert res[1][0][:3] == [[4.0, 5.0], [6.0, 7.0], [8.0, 9.0]] and res[1][0][3] == [102.0, 103.0]
    assert res[0][1] == [3.0] * 4 and res[1][1] == [3.0] * 4


def test_symm_arena_suballocation_is_rank_independent(monkeypatch):
    """The peer-mapped exchange buffers are sub-allocated from symmetric-memory chunks; every rank must arrive at the
    same (chunk, offset) for the same request sequence, or a scattered row would land in the wrong exchange."""
    from posetraj_b200.frame_sharding import SymmArena

    def fake_chunk(self, nbytes):
        size = max(self.CHUNK, (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20))
        self.chunks.append([torch.zeros(0).new_empty(size, dtype=torch.uint8), None, 0])
        self.bytes += size

    monkeypatch.setattr(SymmArena, "_new_chunk", fake_chunk)
    monkeypatch.setattr(SymmArena, "CHUNK", 1 << 20)
    reqs = [300_000, 512, 700_001, 1 << 21, 64, 999_999, 5]
    arenas = [SymmArena.__new__(SymmArena) for _ in range(3)]
    for a in arenas:
        a.device, a.group, a.chunks, a.bytes = "cpu", None, [], 0
    slots = [[a.take(n) for n in reqs] for a in arenas]
    assert slots[0] == slots[1] == slots[2]
    seen = {}
    for (c, off), n in zip(slots[0], reqs):
        assert off % 256 == 0 and off + n <= arenas[0].chunks[c][0].numel()
        for (c2, off2, n2) in seen.get(c, []):
            assert off >= off2 + n2 or off2 >= off + n      # no overlap inside a chunk
        seen.setdefault(c, []).append((c, off, n))
    v = arenas[0].view(slots[0][0], 100, 64)
    assert v.shape == (100, 64) and v.dtype == torch.bfloat16 and getattr(v, "_pt_no_pool", False)


@pytest.mark.parametrize("B,Ft,HW,world", [(2, 5, 45, 3), (2, 25, 144, 8), (1, 14, 45, 2), (2, 3, 7, 3)])
def test_scatter_epilogue_row_formulas(B, Ft, HW, world):
    """The fused exchange (PtGemmArgs.scatter_mode 1 / 2): the destination row each rank's GEMM epilogue computes must
    be a bijection onto the owner's layout and carry the right (batch, frame, pixel) — simulated on the host with the
    same integer formulas as csrc/gemm.cu, for ragged frame and pixel shards."""
    from posetraj_b200.frame_sharding import pixel_shards
    from posetraj_b200.sharding import frame_shards
    fsh, pix = frame_shards(Ft, world), pixel_shards(HW, world)

    def owner(u, shards):
        q = 0
        while q + 1 < len(shards) and u >= shards[q][0] + shards[q][1]:
            q += 1
        return q

    # mode 1: frame layout (b, j in F_r, s) -> pixel layout of the owner of s
    dest = [dict() for _ in range(world)]
    for r, (f0, nf) in enumerate(fsh):
        J, S, kept_off, kept_total = nf, HW, f0, Ft
        for row in range(B * J * S):
            b, rem = divmod(row, J * S)
            j, s = divmod(rem, S)
            q = owner(s, pix)
            drow = (b * kept_total + kept_off + j) * pix[q][1] + (s - pix[q][0])
            assert drow not in dest[q]
            dest[q][drow] = (b, f0 + j, s)
    for q, (p0, npx) in enumerate(pix):
        assert sorted(dest[q]) == list(range(B * Ft * npx))
        for drow, (b, f, s) in dest[q].items():
            assert drow == (b * Ft + f) * npx + (s - p0)
    # mode 2: pixel layout (b, f, p in P_r) -> frame layout of the owner of f
    dest = [dict() for _ in range(world)]
    for r, (p0, npx) in enumerate(pix):
        J, S, kept_off, kept_total = Ft, npx, p0, HW
        for row in range(B * J * S):
            b, rem = divmod(row, J * S)
            j, s = divmod(rem, S)
            q = owner(j, fsh)
            drow = (b * fsh[q][1] + (j - fsh[q][0])) * kept_total + kept_off + s
            assert drow not in dest[q]
            dest[q][drow] = (b, j, p0 + s)
    for q, (f0, nf) in enumerate(fsh):
        assert sorted(dest[q]) == list(range(B * nf * HW))
        for drow, (b, f, s) in dest[q].items():
            assert drow == (b * nf + (f - f0)) * HW + s
