"""Frame-sharded execution of ONE video over several GPUs (SURVEY.md §5 "long-context" row, §8e "Frames").

Spatial sub-blocks (ResnetBlock2D, per-frame GroupNorm, the spatial transformer block, down/up-samplers, conv_in/out,
the conditioning embedding) are independent per frame: rank r runs them on its frames F_r in the usual token layout
("F-layout": rows (b, f in F_r, hw)).  Temporal sub-blocks (TemporalResnetBlock: (3,1,1) convs + 5-D GroupNorm; the
temporal transformer block: attention over frames) are independent per pixel: rank r runs them on ALL frames of its
pixel slice P_r ("P-layout": rows (b, f, hw in P_r)).  Between them the activations are re-sharded by an all-to-all
over NVLink — one each way per SpatioTemporalResBlock and per TransformerSpatioTemporalModel, ~110 per step — with
row-block-copy kernels packing the send order ([dest rank][batch row][...]) and unpacking the receive order.  The 5-D GroupNorm statistics
span all pixels, i.e. all ranks: statistics kernel -> all-reduce of [B, 32, 2] fp64 sums -> normalisation kernel.

Frames and pixels are split raggedly when the world size does not divide them (25 frames on 8 GPUs:
4,3,3,3,3,3,3,3).  The reference's temporal cross-attention indexes its 1-token context by the GLOBAL flattened
position (b*HW + s) mod B; a pixel slice passes its first global pixel and its local width to
the GEMM epilogue (PtGemmArgs.rv_mod / rv_off) so the quirk is reproduced exactly under any slicing.

The CFG + Euler update is elementwise, so every rank updates only its own frames' latents; the final latents are
all-gathered once per call.  With NCCL the whole step — kernels and collectives — is captured into one CUDA graph.

Fused exchange (default on NCCL groups, PT_P2P=0 disables): all four tensors that cross the sharding boundary are
GEMM outputs consumed only by the exchange (spatial conv2 / spatial ff.out going to the pixel layout, temporal conv2 /
temporal ff.out+mix coming back), so the GEMM's epilogue scatters each finished row straight into the owning rank's
buffer over NVLink peer memory (PtGemmArgs.scatter_mode; buffers from torch symmetric memory): pack kernel, NCCL
all-to-all and unpack kernel become ZERO extra kernels, and one device-side barrier per exchange orders the consumers.
Every exchange owns a dedicated destination buffer, so a buffer is only rewritten one whole step later, after >= 100
barriers that its readers have passed: no second ("ready to receive") barrier is needed.  gloo groups (test rigs with
several ranks on one GPU) keep the pack / all-to-all / unpack path.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

from . import ops
from .engine import BF16, F32, NetPlan
from .sharding import frame_shards


def pixel_shards(num_pixels: int, world_size: int) -> List[Tuple[int, int]]:
    """(start, count) of the contiguous pixel slice of every rank (ragged when world_size does not divide HW)."""
    return frame_shards(num_pixels, world_size)


def f2p_tables(B: int, nf: int, HW: int, pix: List[Tuple[int, int]]):
    """Pack tables, frame -> pixel direction: F-layout rows (b, f_local, hw) -> the all-to-all send order
    [dest rank][b][f_local][hw in P_dest].  Returns (src_rows, dst_rows, rows)."""
    src_rows, dst_rows, rows = [], [], []
    off = 0
    for q0, nq in pix:
        for b in range(B):
            for f in range(nf):
                src_rows.append((b * nf + f) * HW + q0)
                dst_rows.append(off + (b * nf + f) * nq)
                rows.append(nq)
        off += B * nf * nq
    return src_rows, dst_rows, rows


def f2p_unpack_tables(B: int, Ft: int, npx: int, fsh: List[Tuple[int, int]]):
    """Unpack tables, frame -> pixel direction: receive order [source rank][b][f in F_source][p_local] -> P-layout
    rows (b, f, p_local)."""
    src_rows, dst_rows, rows = [], [], []
    off = 0
    for f0, cnt in fsh:
        for b in range(B):
            src_rows.append(off + b * cnt * npx)
            dst_rows.append((b * Ft + f0) * npx)
            rows.append(cnt * npx)
        off += B * cnt * npx
    return src_rows, dst_rows, rows


def p2f_pack_tables(B: int, Ft: int, npx: int, fsh: List[Tuple[int, int]]):
    """Pack tables, pixel -> frame direction: P-layout rows (b, f, p_local) -> send order [dest rank][b][f in F_dest][p_local]."""
    src, dst, rows = f2p_unpack_tables(B, Ft, npx, fsh)
    return dst, src, rows


def p2f_tables(B: int, nf: int, HW: int, pix: List[Tuple[int, int]]):
    """Unpack tables, pixel -> frame direction: receive order [source rank][b][f_local][hw in P_source] -> F-layout."""
    src, dst, rows = f2p_tables(B, nf, HW, pix)
    return dst, src, rows


class _Collective:
    """Base of the torch.distributed steps inside an op list (they run on the current stream)."""
    kind, alg_flops, alg_bytes = "collective", 0.0, 0.0


class AllToAllRows(_Collective):
    """all_to_all_single on [rows, C] tensors with per-peer row counts; NCCL on GPUs, host-staged for gloo groups."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, send_rows: List[int], recv_rows: List[int], group, name="a2a"):
        assert src.shape[0] == sum(send_rows) and dst.shape[0] == sum(recv_rows), (src.shape, dst.shape, send_rows, recv_rows)
        self.src, self.dst, self.send_rows, self.recv_rows, self.group, self.name = src, dst, send_rows, recv_rows, group, name
        self.alg_bytes = float(src.numel() * 2)

    def launch(self, stream_ptr: int) -> None:
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all_single(self.dst, self.src, self.recv_rows, self.send_rows, group=self.group)
        else:
            out = torch.empty(self.dst.shape, dtype=self.dst.dtype)
            dist.all_to_all_single(out, self.src.cpu(), self.recv_rows, self.send_rows, group=self.group)
            self.dst.copy_(out)


class AllReduceSum(_Collective):
    def __init__(self, t: torch.Tensor, group, name="all_reduce"):
        self.t, self.group, self.name = t, group, name

    def launch(self, stream_ptr: int) -> None:
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.t, group=self.group)
        else:
            h = self.t.cpu()
            dist.all_reduce(h, group=self.group)
            self.t.copy_(h)


class SymmArena:
    """Peer-mapped exchange buffers: chunks of torch symmetric memory, sub-allocated identically on every rank (all
    ranks build the same plan and ask for the same worst-case sizes in the same order, so offsets agree)."""
    CHUNK = 256 << 20

    def __init__(self, device, group):
        import torch.distributed as dist
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.chunks = []     # [buffer (uint8), handle, used bytes]
        self.bytes = 0

    def _new_chunk(self, nbytes: int) -> None:
        import torch.distributed._symmetric_memory as sm
        size = max(self.CHUNK, (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20))
        try:
            buf = sm.empty(size, dtype=torch.uint8, device=self.device)
            hdl = sm.rendezvous(buf, self.group)
        except Exception as e:  # noqa: BLE001 - say what to do instead of a bare driver error
            raise RuntimeError("posetraj_b200: peer-mapped exchange buffers (torch symmetric memory) are unavailable for "
                               "this process group — all ranks must sit in one NVLink/NVSwitch domain with peer access; "
                               "set PT_P2P=0 to use the NCCL all-to-all exchange instead") from e
        buf.zero_()
        self.chunks.append([buf, hdl, 0])
        self.bytes += size

    def take(self, nbytes: int) -> Tuple[int, int]:
        nbytes = (nbytes + 255) // 256 * 256
        if not self.chunks or self.chunks[-1][2] + nbytes > self.chunks[-1][0].numel():
            self._new_chunk(nbytes)
        c = self.chunks[-1]
        off = c[2]
        c[2] += nbytes
        return len(self.chunks) - 1, off

    def view(self, slot: Tuple[int, int], rows: int, cols: int) -> torch.Tensor:
        buf = self.chunks[slot[0]][0]
        t = buf[slot[1]: slot[1] + rows * cols * 2].view(BF16).view(rows, cols)
        t._pt_no_pool = True
        return t

    def peer_ptrs(self, slot: Tuple[int, int]) -> List[int]:
        return [int(p) + slot[1] for p in self.chunks[slot[0]][1].buffer_ptrs]

    def barrier(self) -> None:
        self.chunks[0][1].barrier(channel=0)


class SymmBarrier(_Collective):
    """Device-side barrier over the symmetric-memory signal pads (no NCCL): every rank's scattered rows have landed."""

    def __init__(self, arena: SymmArena, name="barrier"):
        self.arena, self.name = arena, name

    def launch(self, stream_ptr: int) -> None:
        self.arena.barrier()


class ShardedNetPlan(NetPlan):
    """NetPlan of one rank: spatial geometry = this rank's frames, temporal geometry = this rank's pixel slices."""

    def __init__(self, kind, cfg, weights, *, batch, frames_total, world, rank, group, height, width, device, **kw):
        self.F_total = frames_total
        self.world, self.rank, self.group = world, rank, group
        self.fshards = frame_shards(frames_total, world)
        self.f0, self.nf = self.fshards[rank]
        if self.nf < 1:
            raise ValueError("more ranks than frames")
        self._sums = []
        import torch.distributed as dist
        self.arena: Optional[SymmArena] = kw.pop("arena", None)
        self.p2p = (world > 1 and os.environ.get("PT_P2P", "1") != "0" and dist.is_initialized()
                    and dist.get_backend(group) == "nccl")
        if self.p2p and world > 8:
            raise ValueError("posetraj_b200: the fused peer-memory exchange addresses at most 8 ranks (one NVSwitch "
                             "domain); set PT_P2P=0 for the NCCL exchange on larger groups")
        if self.p2p and self.arena is None:
            self.arena = SymmArena(device, group)
        super().__init__(kind, cfg, weights, batch=batch, frames=self.nf, height=height, width=width, device=device, **kw)

    # ---- layout exchange ------------------------------------------------------------------------------------
    def _pix(self, HW: int):
        sh = pixel_shards(HW, self.world)
        if min(c for _, c in sh) < 1:
            raise ValueError(f"more ranks than pixels at a level with {HW} pixels")
        return sh

    def to_pixel_layout(self, x: torch.Tensor, HW: int, name: str) -> torch.Tensor:
        """F-layout [B*nf*HW, C] -> P-layout [B*F*np, C]: pack -> ONE all-to-all (both batch rows) -> unpack."""
        B, nf, Ft, Cc = self.B, self.nf, self.F_total, x.shape[1]
        sh = self._pix(HW)
        npx = sh[self.rank][1]
        send_buf = self.pool.get(B * nf * HW, Cc)
        recv_buf = self.pool.get(B * Ft * npx, Cc)
        self.step_ops.append(ops.RowBlockCopy(x, send_buf, *f2p_tables(B, nf, HW, sh), name=name + ".pack"))
        self.step_ops.append(AllToAllRows(send_buf, recv_buf, [B * nf * nq for _, nq in sh],
                                          [B * cnt * npx for _, cnt in self.fshards], self.group, name=name + ".f2p"))
        out = self.pool.get(B * Ft * npx, Cc)
        self.step_ops.append(ops.RowBlockCopy(recv_buf, out, *f2p_unpack_tables(B, Ft, npx, self.fshards), name=name + ".unpack"))
        self.pool.put(send_buf, recv_buf)
        return out

    def to_frame_layout(self, y: torch.Tensor, HW: int, name: str) -> torch.Tensor:
        """P-layout [B*F*np, C] -> F-layout [B*nf*HW, C]."""
        B, nf, Ft, Cc = self.B, self.nf, self.F_total, y.shape[1]
        sh = self._pix(HW)
        npx = sh[self.rank][1]
        send_buf = self.pool.get(B * Ft * npx, Cc)
        recv_buf = self.pool.get(B * nf * HW, Cc)
        self.step_ops.append(ops.RowBlockCopy(y, send_buf, *p2f_pack_tables(B, Ft, npx, self.fshards), name=name + ".pack"))
        self.step_ops.append(AllToAllRows(send_buf, recv_buf, [B * cnt * npx for _, cnt in self.fshards],
                                          [B * nf * nq for _, nq in sh], self.group, name=name + ".p2f"))
        out = self.pool.get(B * nf * HW, Cc)
        self.step_ops.append(ops.RowBlockCopy(recv_buf, out, *p2f_tables(B, nf, HW, sh), name=name + ".unpack"))
        self.pool.put(send_buf, recv_buf)
        return out

    def _gemm_exchange(self, direction: str, HW: int, a0, wt, n_cols: int, *, name: str, **kw) -> torch.Tensor:
        """A GEMM whose output only exists to change sharding: 'f2p' (frame layout -> pixel layout) or 'p2f'.
        P2P: the epilogue scatters the rows into the owners' buffers (one kernel) + one barrier.  Otherwise: GEMM into a
        local tensor, then pack / all-to-all / unpack."""
        B, nf, Ft = self.B, self.nf, self.F_total
        sh = self._pix(HW)
        p0, npx = sh[self.rank]
        if not self.p2p:
            tmp = self._gemm(a0, wt, n_cols, name=name, **kw)
            out = (self.to_pixel_layout if direction == "f2p" else self.to_frame_layout)(tmp, HW, name)
            self.pool.put(tmp)
            return out
        if direction == "f2p":
            slot = self.arena.take(B * Ft * max(c for _, c in sh) * n_cols * 2)
            local = self.arena.view(slot, B * Ft * npx, n_cols)
            spec = dict(mode=1, B=B, J=nf, S=HW, kept_off=self.f0, kept_total=Ft, starts=[s for s, _ in sh],
                        counts=[c for _, c in sh], peers=self.arena.peer_ptrs(slot))
        else:
            slot = self.arena.take(B * max(c for _, c in self.fshards) * HW * n_cols * 2)
            local = self.arena.view(slot, B * nf * HW, n_cols)
            spec = dict(mode=2, B=B, J=Ft, S=npx, kept_off=p0, kept_total=HW, starts=[s for s, _ in self.fshards],
                        counts=[c for _, c in self.fshards], peers=self.arena.peer_ptrs(slot))
        self._gemm(a0, wt, n_cols, out=local, scatter=spec, name=name + "+" + direction, **kw)
        self.step_ops.append(SymmBarrier(self.arena, name=name + ".barrier"))
        return local

    def _gn_temporal(self, x: torch.Tensor, key: str, HW: int, npx: int, eps: float) -> torch.Tensor:
        """5-D GroupNorm of a pixel-sharded tensor: statistics span all ranks."""
        Cc = x.shape[1]
        out = self.pool.get(x.shape[0], Cc)
        gam, bet = self.w.f32(key + ".weight"), self.w.f32(key + ".bias")
        common = dict(rows_per_stat=self.F_total * npx, eps=eps, silu=True)
        count = float(Cc // 32) * self.F_total * HW
        if self.p2p:
            # every rank publishes its sums in peer-mapped memory; after one barrier the apply kernel adds the ranks'
            # sums itself, in rank order (no NCCL all-reduce; a dedicated slot per norm, rewritten one step later)
            slot = self.arena.take(self.B * 64 * 8)
            buf = self.arena.chunks[slot[0]][0]
            sums = buf[slot[1]: slot[1] + self.B * 64 * 8].view(torch.float64)
            self._sums.append(sums)
            self.step_ops.append(ops.GroupNorm(x, out, gam, bet, self.stats, mode=1, sums=sums, name=key + ".stats", **common))
            self.step_ops.append(SymmBarrier(self.arena, name=key + ".barrier"))
            self.step_ops.append(ops.GroupNorm(x, out, gam, bet, self.stats, mode=2, sums=sums, count=count,
                                               sums_peers=self.arena.peer_ptrs(slot), name=key + ".apply", **common))
            return out
        sums = torch.zeros(self.B * 64, device=self.device, dtype=torch.float64)
        self._sums.append(sums)
        self.step_ops.append(ops.GroupNorm(x, out, gam, bet, self.stats, mode=1, sums=sums, name=key + ".stats", **common))
        self.step_ops.append(AllReduceSum(sums, self.group, name=key + ".allreduce"))
        self.step_ops.append(ops.GroupNorm(x, out, gam, bet, self.stats, mode=2, sums=sums, count=count, name=key + ".apply", **common))
        return out

    # ---- blocks ---------------------------------------------------------------------------------------------
    def resblock(self, prefix, x0, x1, cout, hw, eps, *, out2=None, aux=None, aux_scale=0.0, res2=None):
        w, B, nf, Ft = self.w, self.B, self.nf, self.F_total
        H, W = hw
        HW = H * W
        rows = self.n * HW
        cin = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
        s, t = prefix + "spatial_res_block.", prefix + "temporal_res_block."
        taps = ops.conv3x3_taps(W)
        # spatial ResnetBlock2D on this rank's frames (identical to the unsharded plan)
        g1 = self._gn(x0, x1, s + "norm1", rows_per_stat=HW, eps=eps, silu=True, halo=hw)
        h1 = self._gemm(g1, w.conv3(s + "conv1.weight"), cout, taps=taps, bias=w.f32(s + "conv1.bias"),
                        rowvec=self._tvec(s, cout), rowvec_mode=1, rv=(nf * HW, 1, 1), halo=hw, out_rows=rows, name=s + "conv1")
        self.pool.put(g1)
        g2 = self._gn(h1, None, s + "norm2", rows_per_stat=HW, eps=eps, silu=True, halo=hw)
        self.pool.put(h1)
        if cin != cout:
            sc = self._gemm(x0, w.linear(s + "conv_shortcut.weight"), cout, a1=x1, bias=w.f32(s + "conv_shortcut.bias"),
                            name=s + "conv_shortcut")
        else:
            sc = x0
        # the spatial block's output goes straight to the pixel layout (fused into the conv2 epilogue with P2P)
        xs_p = self._gemm_exchange("f2p", HW, g2, w.conv3(s + "conv2.weight"), cout, taps=taps, bias=w.f32(s + "conv2.bias"),
                                   res1=sc, halo=hw, out_rows=rows, name=s + "conv2")
        self.pool.put(g2)
        if sc is not x0:
            self.pool.put(sc)
        # TemporalResnetBlock on all frames of this rank's pixel slice
        npx = self._pix(HW)[self.rank][1]
        t1 = self._gn_temporal(xs_p, t + "norm1", HW, npx, eps)
        t2 = self._gemm(t1, w.tconv(t + "conv1.weight"), cout, batches=B, taps=(-npx, 0, npx), bias=w.f32(t + "conv1.bias"),
                        rowvec=self._tvec(t, cout), rowvec_mode=1, rv=(Ft * npx, 1, 1), name=t + "conv1")
        self.pool.put(t1)
        t3 = self._gn_temporal(t2, t + "norm2", HW, npx, eps)
        self.pool.put(t2)
        alpha = w.alpha(prefix + "time_mixer.mix_factor")
        out = self._gemm_exchange("p2f", HW, t3, w.tconv(t + "conv2.weight"), cout, batches=B, taps=(-npx, 0, npx),
                                  bias=w.f32(t + "conv2.bias"), acc_scale=1.0 - alpha, res1=xs_p, name=t + "conv2")
        self.pool.put(t3, xs_p)
        # the residual extras live in the frame layout: mid residual (res2) and ControlNet skip injection (out2/aux)
        if res2 is not None:
            self.step_ops.append(ops.Axpy(out, res2, out, 1.0, name=prefix + "res2"))
        if out2 is not None:
            self.step_ops.append(ops.Axpy(out, aux, out2, aux_scale, name=prefix + "inject"))
        return out

    def _frame_pos_emb(self, prefix: str, Cc: int) -> torch.Tensor:
        saved, self.F = self.F, self.F_total   # the temporal block sees ALL frames
        try:
            return super()._frame_pos_emb(prefix, Cc)
        finally:
            self.F = saved

    def transformer(self, prefix, x, heads, hw, *, out2=None, aux=None, aux_scale=0.0):
        w, B, nf, Ft = self.w, self.B, self.nf, self.F_total
        H, W = hw
        HW = H * W
        Cc = x.shape[1]
        sb, tb = prefix + "transformer_blocks.0.", prefix + "temporal_transformer_blocks.0."
        xvec_s = self._xvec(sb + "attn2.", Cc)[self.row_offset: self.row_offset + B]
        xvec_t = self._xvec(tb + "attn2.", Cc)
        pos = self._frame_pos_emb(prefix, Cc)
        g = self._gn(x, None, prefix + "norm", rows_per_stat=HW, eps=1e-6, silu=False)
        h = self._gemm(g, w.linear(prefix + "proj_in.weight"), Cc, bias=w.f32(prefix + "proj_in.bias"), name=prefix + "proj_in")
        self.pool.put(g)
        # --- spatial BasicTransformerBlock on this rank's frames
        l1 = self._ln(h, sb + "norm1")
        qkv = self._gemm(l1, w.qkv(sb + "attn1."), 3 * Cc, name=sb + "attn1.qkv")
        self.pool.put(l1)
        att = self.pool.get(x.shape[0], Cc)
        self.step_ops.append(ops.AttnSpatial(qkv, att, n_img=self.n, heads=heads, name=sb + "attn1"))
        self.pool.put(qkv)
        h2 = self._gemm(att, w.linear(sb + "attn1.to_out.0.weight"), Cc, bias=w.f32(sb + "attn1.to_out.0.bias"), res1=h,
                        rowvec=xvec_s, rowvec_mode=1, rv=(nf * HW, 1, 1), name=sb + "attn1.to_out")
        self.pool.put(att, h)
        l3 = self._ln(h2, sb + "norm3")
        f1 = self._gemm(l3, w.linear(sb + "ff.net.0.proj.weight"), 4 * Cc, geglu=True, bias=w.f32(sb + "ff.net.0.proj.bias"),
                        name=sb + "ff.geglu")
        self.pool.put(l3)
        h3_p = self._gemm_exchange("f2p", HW, f1, w.linear(sb + "ff.net.2.weight"), Cc, bias=w.f32(sb + "ff.net.2.bias"),
                                   res1=h2, name=sb + "ff.out")
        self.pool.put(f1, h2)
        # --- TemporalBasicTransformerBlock on all frames of this rank's pixel slice
        p0, npx = self._pix(HW)[self.rank]
        rows_p = h3_p.shape[0]
        ht = self.pool.get(rows_p, Cc)
        l_in = self._ln(h3_p, tb + "norm_in", addvec=pos, hw=npx, frames=Ft, sum_out=ht)
        fi = self._gemm(l_in, w.linear(tb + "ff_in.net.0.proj.weight"), 4 * Cc, geglu=True,
                        bias=w.f32(tb + "ff_in.net.0.proj.bias"), name=tb + "ff_in.geglu")
        self.pool.put(l_in)
        t1 = self._gemm(fi, w.linear(tb + "ff_in.net.2.weight"), Cc, bias=w.f32(tb + "ff_in.net.2.bias"), res1=ht,
                        name=tb + "ff_in.out")
        self.pool.put(fi, ht)
        l1t = self._ln(t1, tb + "norm1")
        qkv_t = self._gemm(l1t, w.qkv(tb + "attn1."), 3 * Cc, name=tb + "attn1.qkv")
        self.pool.put(l1t)
        att_t = self.pool.get(rows_p, Cc)
        self.step_ops.append(ops.AttnTemporal(qkv_t, att_t, batch=B, frames=Ft, hw=npx, heads=heads, name=tb + "attn1"))
        self.pool.put(qkv_t)
        # context of hidden row (b, s): batch ((b*HW + s) mod ctx_B) with the GLOBAL pixel index s = p0 + local
        t2 = self._gemm(att_t, w.linear(tb + "attn1.to_out.0.weight"), Cc, bias=w.f32(tb + "attn1.to_out.0.bias"), res1=t1,
                        rowvec=xvec_t, rowvec_mode=2, rv=(Ft * npx, HW, self.ctx_B, npx, p0 + self.row_offset * HW),
                        name=tb + "attn1.to_out")
        self.pool.put(att_t, t1)
        l3t = self._ln(t2, tb + "norm3")
        f2 = self._gemm(l3t, w.linear(tb + "ff.net.0.proj.weight"), 4 * Cc, geglu=True, bias=w.f32(tb + "ff.net.0.proj.bias"),
                        name=tb + "ff.geglu")
        self.pool.put(l3t)
        alpha = w.alpha(prefix + "time_mixer.mix_factor")
        hb = self._gemm_exchange("p2f", HW, f2, w.linear(tb + "ff.net.2.weight"), Cc, bias=w.f32(tb + "ff.net.2.bias"),
                                 acc_scale=1.0 - alpha, res1=t2, res1_scale=1.0 - alpha, res2=h3_p, res2_scale=alpha,
                                 name=tb + "ff.out+mix")
        self.pool.put(f2, t2, h3_p)
        out = self._gemm(hb, w.linear(prefix + "proj_out.weight"), Cc, bias=w.f32(prefix + "proj_out.bias"), res1=x,
                         out2=out2, aux=aux, aux_scale=aux_scale, name=prefix + "proj_out")
        self.pool.put(hb)
        return out


class FrameShardedEngine:
    """One video's denoise loop on `world` ranks, this rank owning frames [f0, f0 + nf)."""

    def __init__(self, unet, controlnet, scheduler, *, frames: int, h: int, w: int, cond_hw: tuple, device, rank: int,
                 world: int, group=None):
        from .engine import WeightStore  # noqa: F401  (weights are shared with the unsharded plans)
        self.unet, self.controlnet, self.scheduler = unet, controlnet, scheduler
        self.F, self.h, self.w, self.device = frames, h, w, device
        self.rank, self.world, self.group = rank, world, group
        self.f0, self.nf = frame_shards(frames, world)[rank]
        cfg = unet.cfg
        self.latents = torch.zeros(self.nf, cfg.out_channels, h, w, device=device, dtype=F32)
        self.image_latents = torch.zeros(2, self.nf, cfg.out_channels, h, w, device=device, dtype=F32)
        self.guidance = torch.ones(self.nf, device=device, dtype=F32)
        self.step_index = torch.zeros(1, device=device, dtype=torch.int32)
        self.sigmas = torch.zeros(1024, device=device, dtype=F32)
        import torch.distributed as dist
        self.arena = None
        if world > 1 and os.environ.get("PT_P2P", "1") != "0" and dist.is_initialized() and dist.get_backend(group) == "nccl":
            self.arena = SymmArena(device, group)    # one set of peer-mapped exchange buffers for both networks
        common = dict(batch=2, frames_total=frames, world=world, rank=rank, group=group, height=h, width=w, device=device,
                      sigmas=self.sigmas, step_index=self.step_index, arena=self.arena)
        self.cplan = ShardedNetPlan("controlnet", controlnet.cfg, controlnet.weights, cond_hw=cond_hw, **controlnet.flags, **common)
        self.uplan = ShardedNetPlan("unet", unet.cfg, unet.weights, x_in=self.cplan.x_in, residual_bufs=self.cplan.res, **common)
        kw = dict(latents=self.latents, guidance=self.guidance, sigmas=self.sigmas, step_index=self.step_index,
                  next_in=self.cplan.x_in, image_latents=self.image_latents, next_padded=True)
        self.prepare_op = ops.CfgEuler(noise_pred=None, mode=1, **kw)
        self.update_op = ops.CfgEuler(noise_pred=self.uplan.noise_pred, mode=0, **kw)
        self.step_ops = self.cplan.step_ops + self.uplan.step_ops + [self.update_op, ops.StepAdvance(self.step_index)]
        self.graph = None
        self.launches_per_step = sum(1 for o in self.step_ops if not isinstance(o, _Collective))
        self.collectives_per_step = sum(1 for o in self.step_ops if isinstance(o, _Collective))

    def load(self, *, latents, image_latents, image_embeddings, added_time_ids, guidance, sigmas, controlnet_condition,
             camera_cond=None, cond_scale: float = 1.0) -> None:
        sp = torch.cuda.current_stream().cuda_stream
        f0, nf = self.f0, self.nf
        lat = latents.reshape(self.F, *self.latents.shape[1:])
        self.latents.copy_(lat[f0:f0 + nf])
        il = image_latents.reshape(2, self.F, *self.latents.shape[1:])
        self.image_latents.copy_(il[:, f0:f0 + nf])
        self.guidance.copy_(guidance.reshape(-1)[f0:f0 + nf])
        self.sigmas[:sigmas.numel()].copy_(sigmas)
        self.step_index.zero_()
        for plan in (self.cplan, self.uplan):
            plan.ehs.copy_(image_embeddings[:, 0, :])
            plan.time_ids.copy_(added_time_ids.reshape(-1))
            NetPlan.run(plan.embed_ops, sp)
        cond = controlnet_condition[:, f0:f0 + nf].contiguous()
        cam = None if camera_cond is None else camera_cond[:, f0:f0 + nf].contiguous()
        self.controlnet.stage_condition(self.cplan, cond, cam, None, sp)
        self.cplan.set_conditioning_scale(cond_scale)
        self.prepare_op.launch(sp)
        self._latents0 = self.latents.clone()

    def reset(self) -> None:
        self.latents.copy_(self._latents0)
        self.step_index.zero_()
        self.prepare_op.launch(torch.cuda.current_stream().cuda_stream)

    def capture(self) -> None:
        """One CUDA graph of the whole step INCLUDING the NCCL all-to-alls / all-reduces (NCCL collectives are
        capturable): removes ~1250 host-side launches per step.  gloo groups (host-staged exchange) stay eager."""
        import torch.distributed as dist
        if self.graph is not None or dist.get_backend(self.group) != "nccl":
            return
        saved = self.step_index.clone(), self.latents.clone(), self.cplan.x_in.clone()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                NetPlan.run(self.step_ops, torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().wait_stream(s)
        self.step_index.copy_(saved[0]); self.latents.copy_(saved[1]); self.cplan.x_in.copy_(saved[2])
        self.graph = g

    def step(self, use_graph: bool = True) -> None:
        if use_graph and self.graph is not None:
            self.graph.replay()
        else:
            NetPlan.run(self.step_ops, torch.cuda.current_stream().cuda_stream)

    def gather_latents(self) -> torch.Tensor:
        """All frames' latents [F, C, h, w] on every rank (ragged frame shards are padded for the all-gather)."""
        import torch.distributed as dist
        shards = frame_shards(self.F, self.world)
        mx = max(c for _, c in shards)
        pad = torch.zeros(mx, *self.latents.shape[1:], device=self.device, dtype=F32)
        pad[:self.nf] = self.latents
        if dist.get_backend(self.group) != "nccl":
            pad = pad.cpu()
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat([p[:c] for p, (_, c) in zip(parts, shards)], 0).to(self.device)
