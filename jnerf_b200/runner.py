"""Thin training / evaluation driver: the caller of the hot path, mirroring runner/runner.py:14-264 step for step
(SURVEY.md section 3.1).  `train_step` is the B200 fast path: no autograd graph, no host sync (except the
reference's own batch-size adaptation every 16 steps), one C-ABI call per stage:

    [grid update /16] -> raygen -> march -> (compaction bookkeeping) -> fused network fwd ->
    fused composite + Huber + composite bwd -> fused network bwd -> [NCCL all-reduce] -> fused Adam+EMA

On one GPU the steps are software-pipelined: raygen + march of step i+1 run on a second stream under step i's Adam+EMA sweep
(`_train_step_pipe`; NGP_PIPELINE=0 restores the strictly sequential step, NGP_GRAPHS=1 the CUDA-graph replay of it).

`train_step_autograd` runs the same step through the per-operator plugin classes and torch autograd, exactly as
JNeRF's Runner.train does with Jittor; tests check that both give the same parameters."""
import contextlib
import os

import numpy as np
import torch

from . import dp, ops
from .plugin import losses as L
from .utils.config import get_cfg
from .utils.registry import DATASETS, LOSSES, NETWORKS, OPTIMS, SAMPLERS, build_from_cfg


class Runner:
    def __init__(self, rank=0, world_size=1, process_group=None):
        self.cfg = get_cfg()
        cfg = self.cfg
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        self.dataset = {"train": build_from_cfg(cfg.dataset.train, DATASETS)}
        cfg.dataset_obj = self.dataset["train"]
        self.dataset["val"] = build_from_cfg(cfg.dataset.val, DATASETS) if cfg.dataset.val else self.dataset["train"]
        self.dataset["test"] = None
        self.model = build_from_cfg(cfg.model, NETWORKS)
        cfg.model_obj = self.model
        self.sampler = build_from_cfg(cfg.sampler, SAMPLERS)
        cfg.sampler_obj = self.sampler
        if world_size > 1:
            self.sampler.dp_group = (process_group, world_size)
        params = list(self.model.parameters())
        self.optimizer = build_from_cfg(cfg.optim, OPTIMS, params=params)
        self.optimizer = build_from_cfg(cfg.expdecay, OPTIMS, nested_optimizer=self.optimizer)
        self.ema_optimizer = build_from_cfg(cfg.ema, OPTIMS, params=params)
        self.ema_optimizer.attach(self.optimizer)
        self.loss_func = build_from_cfg(cfg.loss, LOSSES)
        self.background_color = cfg.background_color
        self.tot_train_steps = cfg.tot_train_steps
        self.n_rays_per_batch = cfg.n_rays_per_batch
        self.W, self.H = self.dataset["train"].resolution
        self.start = 0
        cfg.m_training_step = 0
        self.val_freq = 4096
        self.fast = bool(getattr(self.model, "fused", False))
        # state the evaluation / checkpoint code reads on EVERY kind of model (fused or the nn.Linear fallback)
        self._table_work, self._pending_epoch = None, None
        self._host_stage = None
        self._dev_state = None                                       # device-resident step state + CUDA graphs (single-GPU fast path)
        self._pipe = None                                            # march of step i+1 under the optimizer sweep of step i (single GPU)
        self._st = {id(s.p): s for s in self.optimizer._nested_optimizer.state}
        if world_size > 1 and not self.fast:
            raise ValueError("data-parallel training needs the fused model (fp16=True, use_fully=True): the per-operator autograd step "
                             "shards the ray batch but has no gradient exchange, the replicas would silently diverge")
        if self.fast:
            self._init_fast_path()
        self._bg_gen = torch.Generator(device="cuda").manual_seed(int(cfg.seed or 1) + 7)

    # ------------------------------------------------------------------------------------------ fast path
    def _init_fast_path(self):
        dev = "cuda"
        m = self.model
        cap = self.sampler.target_batch_size
        self.grid_grad = torch.zeros(m.pos_encoder.m_grid.numel(), dtype=torch.float16, device=dev)
        n_w = m.density_mlp.con_weights.numel() + m.rgb_mlp.con_weights.numel()
        # [dW density | dW colour | measured-sample counter]: one flat fp32 buffer = one small all-reduce
        self.w_grad = torch.zeros(n_w + 8, dtype=torch.float32, device=dev)
        self.dwd = self.w_grad[:m.density_mlp.con_weights.numel()]
        self.dwr = self.w_grad[m.density_mlp.con_weights.numel():n_w]
        self.net_out = torch.zeros((cap, 4), dtype=torch.float16, device=dev)
        self.enc = torch.empty((cap, 32), dtype=torch.float16, device=dev)
        self.dnet = torch.zeros((cap, 4), dtype=torch.float16, device=dev)
        self.last_loss = None
        self.last_rgb = None
        if self.world_size == 1:
            # Device-resident step state (include/ngp_b200.h: ngp_step_state_*): the sampler rng, the pixel cursor and Adam's step
            # factors live on the device, so every launch of a training step has the same arguments and the step can be captured in
            # a CUDA graph per ray-batch size.  Opt-in (NGP_GRAPHS=1): measured +2.5 % at 2^18 samples per iteration once every
            # ray-batch size of the run has its graph, nothing over 1 000 steps of a still-converging scene (17 sizes, 17 captures),
            # and -4 % at 2^20 where a capture allocates hundreds of MB (profiles/r02_kernels/call17, r02_final).
            self._dev_state = ops.step_state_new()
            self._dev_expect = None
            self._graphs, self._graph_seen, self._graph_pool, self._cap_stream = {}, {}, None, None
            self._graphs_enabled = os.environ.get("NGP_GRAPHS", "0") == "1" and torch.cuda.is_available() and hasattr(torch.cuda, "CUDAGraph")
            self.graph_replays = 0
            self._graph_after = int(os.environ.get("NGP_GRAPH_AFTER", "20"))     # occurrences of a ray-batch size before it gets a graph
        # Software pipeline over steps (default): nothing the march reads is written by the network kernels -- rays, jitter and the
        # occupancy bitfield only -- so the "front" of step i+1 (background colours, ray generation, march, compaction) is enqueued
        # on a second stream while step i's backward / Adam+EMA sweep still run.  The sweep is HBM-bound and the march is
        # latency-bound: side by side they share the SMs instead of queueing.  NGP_PIPE_AT picks the point of
        # step i the front of step i+1 may start at: after its network forward ("fwd"), backward ("bwd") or at once ("front").
        if os.environ.get("NGP_PIPELINE", "1") == "1" and not getattr(self, "_graphs_enabled", False):
            at = os.environ.get("NGP_PIPE_AT", "front")
            assert at in ("front", "fwd", "bwd")
            # both coordinate buffers are allocated (and zero-filled) HERE, on the constructor's stream: a fill enqueued lazily from inside
            # a step would land on the main stream behind that step's kernels and wipe what the side stream's march had just written
            raw = self.sampler._coords_raw
            self._pipe = dict(stream=torch.cuda.Stream(), coords=[torch.zeros_like(raw), torch.zeros_like(raw)], made=0, pending=None, at=at, mid=torch.cuda.Event(),
                              back_done=[torch.cuda.Event(), torch.cuda.Event()], prefetched=0, aux=torch.cuda.Stream(),
                              bwd_done=torch.cuda.Event(), aux_done=torch.cuda.Event(), main=None)
        if self.world_size > 1:
            self._init_sharded_table()

    def _init_sharded_table(self):
        """Sharded optimizer for the hash table (dp.py header): this rank owns slice [lo, hi) of the padded table.
        Exchange mode (env NGP_DP_EXCHANGE = auto | p2p | nccl):
          p2p  -- ONE kernel per step over NVLink peer memory (ngp_dp_exchange_step): pull-reduce the peers' gradient slices,
                  Adam+EMA on the slice, push the fp16 slice into every peer's table, all-reduce + update the MLP weights;
          nccl -- reduce-scatter -> Adam+EMA on the slice -> async all-gather, + all-reduce of the MLP weight gradients.
        Either way the table of the next step is awaited only after the march (which does not read it)."""
        import os
        g = self.model.pos_encoder.m_grid
        m = self.model
        n, W = g.numel(), self.world_size
        mode = os.environ.get("NGP_DP_EXCHANGE", "auto")
        assert mode in ("auto", "p2p", "nccl")
        self.dp_mode = "p2p" if mode == "p2p" or (mode == "auto" and dp.peer_exchange_available(W)) else "nccl"
        P = dp.padded_len(n, W)
        nd, nr = m.density_mlp.con_weights.numel(), m.rgb_mlp.con_weights.numel()
        if self.dp_mode == "p2p":
            try:
                self._arena = dp.PeerArena(n, nd + nr, W, self.rank, self.pg)
            except RuntimeError as e:                                  # raised on EVERY rank when any mapping failed
                if mode == "p2p":
                    raise
                if self.rank == 0:
                    print(f"[jnerf_b200] peer-memory exchange unavailable ({e}); using the NCCL exchange", flush=True)
                self.dp_mode = "nccl"
        if self.dp_mode == "p2p":
            self._table, self._table_grad, self.w_grad = self._arena.table, self._arena.table_grad, self._arena.w_grad
            self.dwd, self.dwr = self.w_grad[:nd], self.w_grad[nd:nd + nr]
            self._peers = {k: self._arena.peers(k) for k in ("table", "table_grad", "w_grad", "flags")}
            # the two MLP weight tensors and their optimizer state become views of flat buffers (one sweep inside the kernel)
            self._w_param = torch.cat([m.density_mlp.con_weights.data.reshape(-1), m.rgb_mlp.con_weights.data.reshape(-1)]).contiguous()
            m.density_mlp.con_weights.data = self._w_param[:nd].view_as(m.density_mlp.con_weights.data)
            m.rgb_mlp.con_weights.data = self._w_param[nd:].view_as(m.rgb_mlp.con_weights.data)
            std, str_ = self._st[id(m.density_mlp.con_weights)], self._st[id(m.rgb_mlp.con_weights)]
            self._w_m, self._w_v = torch.cat([std.m, str_.m]), torch.cat([std.v, str_.v])
            self._w_master = torch.cat([std.master, str_.master])
            for st, lo_, hi_ in ((std, 0, nd), (str_, nd, nd + nr)):
                st.m, st.v, st.master = self._w_m[lo_:hi_], self._w_v[lo_:hi_], self._w_master[lo_:hi_]
            self._epoch, self._pending_epoch = 0, None
        else:
            self._table = torch.zeros(P, dtype=g.dtype, device=g.device)
            self._table_grad = torch.zeros(P, dtype=self.grid_grad.dtype, device=g.device)
        self._table[:n].copy_(g.data.reshape(-1))
        g.data = self._table[:n]                                   # the live parameter is a view of the padded table
        self.grid_grad = self._table_grad[:n]
        self._lo, self._hi = dp.slice_bounds(P, W, self.rank)
        self._slice_grad = torch.empty(self._hi - self._lo, dtype=self.grid_grad.dtype, device=g.device)
        self._slice_param = self._table[self._lo:self._hi].clone()
        st = self._st[id(g)]
        master = torch.zeros(P, dtype=torch.float32, device=g.device)
        master[:n].copy_(st.master)
        st.master = master[self._lo:self._hi].clone()
        st.m = torch.zeros(self._hi - self._lo, dtype=torch.float32, device=g.device)
        st.v = torch.zeros_like(st.m)

    def _table_ready(self, zero_on=None):
        """Make the current stream wait for the in-flight all-gather of the updated table (no-op on one GPU).
        zero_on: (stream, event) -- clear the gradient buffers on that stream instead of the current one and record the event; the
        caller waits for it before its backward.  The 25 MB memset then runs beside the network forward (which does not touch the
        gradients) instead of in front of it."""
        if self._table_work is not None:
            self._table_work.wait()
            self._table_work = None
        if getattr(self, "_pending_epoch", None) is not None:
            ops.dp_exchange_wait(self.world_size, self._arena.flags, self._pending_epoch)
            self._pending_epoch = None
            # peers no longer read the gradients: clear table + MLP gradients in one memset
            if zero_on is None:
                self._arena.grads.zero_()
                return False
            stream, ev = zero_on
            ev.record(torch.cuda.current_stream())                   # behind the wait kernel
            stream.wait_event(ev)
            with torch.cuda.stream(stream):
                self._arena.grads.zero_()
                ev.record(stream)
            return True
        return False

    def _gather_table_state(self):
        """Full-length (m, v, master) of the hash table for checkpoints: all-gather of the per-rank slices."""
        g = self.model.pos_encoder.m_grid
        st, n = self._st[id(g)], g.numel()
        out = []
        for t in (st.m, st.v, st.master):
            full = torch.empty(self._table.numel(), dtype=torch.float32, device=g.device)
            dp.all_gather_slices(full, t, self.pg, self.world_size)
            out.append(full[:n].clone())
        return out

    def _scatter_table_state(self, m, v, master):
        g = self.model.pos_encoder.m_grid
        st, n = self._st[id(g)], g.numel()
        for dst, src in ((st.m, m), (st.v, v), (st.master, master)):
            full = torch.zeros(self._table.numel(), dtype=torch.float32, device=g.device)
            full[:n].copy_(src.reshape(-1))
            dst.copy_(full[self._lo:self._hi])

    def next_batch(self):
        """Global pixel batch of this step (identical on every rank), sharded contiguously by rank."""
        ds = self.dataset["train"]
        n = self.sampler.n_rays_per_batch                       # per-rank rays; the global batch is world_size x n (weak scaling)
        pix = ds.next_pixels(n * self.world_size)
        if self.world_size > 1:
            pix = pix[self.rank * n:(self.rank + 1) * n]
        img_ids, rays_o, rays_d = ds.rays_for(pix)
        return img_ids, rays_o, rays_d, ds.rgba_for(pix)

    # ------------------------------------------------------------------------------------------ device-state / CUDA-graph step
    def _lr_for_step(self, k):
        """ExpDecay's learning rate for its k-th call (optims/expdecay.py:20-25), without advancing it."""
        dec = self.optimizer
        f = dec.m_learning_rate_factor
        for q in range(dec.steps, k + 1):
            if q >= dec.decay_start and (q - dec.decay_start) % dec.decay_interval == 0 and q <= dec.decay_end:
                f *= dec.decay_base
        return dec.base_lr * f

    def _step_body(self, R, lr_next):
        """Every launch of one training step on R device-generated rays, with nothing step-dependent among the launch arguments (the
        rng, the pixel cursor and Adam's factors come from the device step state): run eagerly or captured into a CUDA graph."""
        s, m, ds, st = self.sampler, self.model, self.dataset["train"], self._dev_state
        adam = self.optimizer._nested_optimizer
        bg = torch.rand((R, 3), device="cuda", generator=self._bg_gen)                               # runner.py:66
        img_ids, rays_o, rays_d, target = ops.prepare_batch_dev(R, ds.shuffle_index, st, 0, ds.W, ds.H, ds.transforms_gpu, ds.focal_lengths,
                                                                ds.principal, ds.image_data, bg)      # dataset.py:172-188 + runner.py:68
        s.sample_dev(rays_o, rays_d, st)                                                              # march, compaction bookkeeping
        coords, n_dev = s.coords_compacted, s.n_samples_dev
        self.net_forward(coords, n_dev)
        rgb, loss, _ = ops.composite_loss_bwd(self.net_out, coords, s._rays_numsteps, s._rays_numsteps_compacted, bg, target,
                                              s.density_grid_mean, delta=self.loss_func.delta, cascades=s.NERF_CASCADES, dnet=self.dnet)
        self.net_backward(coords, n_dev)
        for p, g in ((m.pos_encoder.m_grid, self.grid_grad), (m.density_mlp.con_weights, self.dwd), (m.rgb_mlp.con_weights, self.dwr)):
            stp = self._st[id(p)]
            ops.adam_ema_dev(p.data, g, stp.m, stp.v, stp.master, st, zero_grad=True)
        ops.step_state_tick(st, R, lr_next, adam.betas[0], adam.betas[1], adam.eps, self.ema_optimizer.decay, 1.0)
        self.last_loss, self.last_rgb = loss, rgb
        return loss

    def _train_step_dev(self):
        cfg, s, ds = self.cfg, self.sampler, self.dataset["train"]
        adam, dec = self.optimizer._nested_optimizer, self.optimizer
        i = cfg.m_training_step
        R = s.n_rays_per_batch
        edge = i % s.update_den_freq == 0 or i % s.update_den_freq == s.update_den_freq - 1
        if i % s.update_den_freq == 0:
            s.update_density_grid()                                  # evaluates the density network, advances the sampler rng
        if R > s._march_ws_rays:
            s._ensure_march_ws(2 * R)
            self._graphs.clear()                                     # the captured launches hold the old workspace's address
            self._graph_pool = None                                  # (the allocator drops a pool together with its last graph)
        start = ds.reserve_pixels(R)
        lr, lr_next = self._lr_for_step(dec.steps), self._lr_for_step(dec.steps + 1)
        want = (int(s.rng[0]), int(s.rng[1]), start, adam.n_step, lr)
        if self._dev_expect != want:                                 # first step, or the host changed something outside a step
            ops.step_state_set(self._dev_state, s.rng, start, adam.n_step, lr, adam.betas[0], adam.betas[1], adam.eps, self.ema_optimizer.decay, 1.0)
        key = (R, lr_next)
        g = self._graphs.get(key) if self._graphs_enabled and not edge else None
        if g is not None:
            g[0].replay()
            ops.lib.launch_count += g[1]
            self.graph_replays += 1
            loss = g[2]
            self.last_loss, self.last_rgb = g[2], g[3]
            s._rays_numsteps, s._rays_numsteps_compacted, s._counters_compacted, s._coords = g[4]
        else:
            seen = self._graph_seen.get(key, 0) + 1
            self._graph_seen[key] = seen
            # a capture costs about as much as 200 replays save: only ray-batch sizes that keep coming back (a second 16-step window)
            # get a graph
            if self._graphs_enabled and not edge and seen >= self._graph_after and len(self._graphs) < 64:
                # capture on a side stream with the raw begin / end calls: torch.cuda.graph() would synchronise the device, run the
                # Python garbage collector and empty the allocator cache on every capture (~10 ms each)
                graph = torch.cuda.CUDAGraph()
                if hasattr(graph, "register_generator_state"):
                    graph.register_generator_state(self._bg_gen)
                n0 = ops.lib.launch_count
                if self._cap_stream is None:
                    self._cap_stream = torch.cuda.Stream()
                main = torch.cuda.current_stream()
                self._cap_stream.wait_stream(main)
                with torch.cuda.stream(self._cap_stream):
                    if self._graph_pool is None:
                        graph.capture_begin(capture_error_mode="thread_local")
                    else:
                        graph.capture_begin(pool=self._graph_pool, capture_error_mode="thread_local")
                    loss = self._step_body(R, lr_next)
                    graph.capture_end()
                main.wait_stream(self._cap_stream)
                if self._graph_pool is None:
                    self._graph_pool = graph.pool()
                self._graphs[key] = (graph, ops.lib.launch_count - n0, loss, self.last_rgb,
                                     (s._rays_numsteps, s._rays_numsteps_compacted, s._counters_compacted, s._coords))
                ops.lib.launch_count = n0
                graph.replay()                                       # capture records the launches, the replay performs the step
                ops.lib.launch_count += self._graphs[key][1]
                self.graph_replays += 1
            else:
                loss = self._step_body(R, lr_next)
        # host mirrors of what the step advanced on the device
        ops.pcg32_advance(s.rng)                                     # rng.advance(), ray_sampler.py:61
        dec.advance_lr()
        adam.n_step += 1
        self.ema_optimizer.steps += 1
        self._dev_expect = (int(s.rng[0]), int(s.rng[1]), start + R, adam.n_step, lr_next)
        if i % s.update_den_freq == s.update_den_freq - 1:
            s.update_batch_rays()
        cfg.m_training_step = i + 1
        return loss

    # ------------------------------------------------------------------------------------------ software pipeline over steps
    def _front(self, step, batch, prefetch):
        """The part of training step `step` that does not read a network parameter: background colours, ray generation + target
        lookup (or the blend of a fed batch), march, compaction.  prefetch=True runs it on the side stream, behind the point of the
        step in flight that NGP_PIPE_AT names; otherwise on the current stream (first step, and every step that starts with an
        occupancy-grid update: the update evaluates the density network, so it needs the finished optimizer sweep)."""
        P, s, ds = self._pipe, self.sampler, self.dataset["train"]
        slot = P["made"] & 1
        main = P["main"] if P["main"] is not None else torch.cuda.current_stream()
        side = P["stream"]
        if prefetch:
            side.wait_event(P["mid"])                                # recorded on the main stream inside the step in flight
        # the step that read this slot's coordinate rows two fronts ago must be through (it is, unless the host runs far ahead)
        (side if prefetch else main).wait_event(P["back_done"][slot])
        with (torch.cuda.stream(side) if prefetch else contextlib.nullcontext()):
            rng_before = s.rng.copy()
            W = self.world_size
            if step % s.update_den_freq == 0:
                assert not prefetch
                self._table_ready()                                  # the update evaluates the density network on the exchanged table
                s.update_density_grid()
            if batch is None:
                R = s.n_rays_per_batch                               # per-rank rays; the global batch is W x R (weak scaling)
                pix = ds.next_pixels(R * W)
                bg = torch.rand((R * W, 3), device="cuda", generator=self._bg_gen)                     # runner.py:66 (global batch, then this rank's rows)
                if W > 1:
                    lo, hi = dp.shard_range(R, self.rank)
                    pix, bg = pix[lo:hi], bg[lo:hi]
                bg = bg.contiguous()
                img_ids, rays_o, rays_d, target = ops.prepare_batch(pix.contiguous(), ds.W, ds.H, ds.transforms_gpu, ds.focal_lengths,
                                                                    ds.principal, ds.image_data, bg)    # dataset.py:172-188 + runner.py:68
            else:
                img_ids, rays_o, rays_d, rgba = batch
                R = rays_o.shape[0]
                bg = torch.rand((R * W, 3), device="cuda", generator=self._bg_gen)
                if W > 1:
                    bg = bg[self.rank * R:(self.rank + 1) * R]
                bg = bg.contiguous()
                target = ops.blend_target(rgba.contiguous(), bg)                                       # runner.py:68
            numsteps, ns_c, cnt_c, coords = s.sample_front(rays_o, rays_d, P["coords"][slot], ray_index_offset=dp.shard_range(R, self.rank)[0])
            done = torch.cuda.Event()
            done.record(side if prefetch else main)
        P["made"] += 1
        P["prefetched"] += int(prefetch)
        return dict(step=step, src=batch, slot=slot, bg=bg, target=target, numsteps=numsteps, ns_c=ns_c, cnt_c=cnt_c, coords=coords, done=done,
                    rng_before=rng_before, rays=(img_ids, rays_o, rays_d), side=prefetch)

    def _adapt_ray_batch(self, F):
        """DensityGridSampler.update_batch_rays (density_grid_sampler.py:266-271) without stalling the pipeline: the 16-step sample
        counter is read back on the SIDE stream (all-reduced there first under data parallelism), behind the last march that added
        to it -- the host waits for that march only, not for the network kernels of the step it has just enqueued, and keeps its
        lead over the device (a `.item()` on the main stream drained the queue every 16 steps: the device then idled while the host
        enqueued the next step from scratch)."""
        P, s = self._pipe, self.sampler
        side = P["stream"]
        if P.get("count_host") is None:
            P["count_host"], P["count_ev"] = torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event()
        side.wait_event(F["done"])
        with torch.cuda.stream(side):
            if self.world_size > 1:
                dp.global_sum_count(s.measured_batch_size, self.pg, self.world_size)
            P["count_host"].copy_(s.measured_batch_size, non_blocking=True)
            s.measured_batch_size.zero_()
            P["count_ev"].record(side)
        P["count_ev"].synchronize()
        P["main"].wait_event(P["count_ev"])                          # the next front (a grid-update step: main stream) adds to the counter
        s.update_batch_rays(measured_total=int(P["count_host"][0]))

    def _sync_front(self):
        """Evaluation and checkpoint code shares the march workspace with a prefetched front: order the current stream behind it."""
        P = self._pipe
        if P is not None and P["pending"] is not None:
            torch.cuda.current_stream().wait_event(P["pending"]["done"])

    def _train_step_pipe(self, batch=None, next_batch=None):
        cfg, s, P = self.cfg, self.sampler, self._pipe
        i = cfg.m_training_step
        main = P["main"] = torch.cuda.current_stream()
        F, P["pending"] = P["pending"], None
        if F is not None and (F["step"] != i or F["src"] is not batch):
            main.wait_event(F["done"])                               # it may still be marching: the workspace is shared with the new front
            F = None                                                 # the caller changed course (other batch, other step): drop it
        if F is None:
            F = self._front(i, batch, prefetch=False)
        elif F["side"]:
            main.wait_event(F["done"])
        if P["at"] == "front":
            P["mid"].record(main)
        # (the sampler is an nn.Module: plain attribute assignment goes through Module.__setattr__, ~2 us apiece)
        s.__dict__.update(_rays_numsteps=F["numsteps"], _rays_numsteps_compacted=F["ns_c"], _counters_compacted=F["cnt_c"], _coords=F["coords"])
        coords, n_dev = F["coords"], F["cnt_c"][0:1]
        # data parallel: the exchange of step i-1 has delivered the table; its gradient buffers are cleared beside the forward
        zeroing = self._table_ready(zero_on=(P["aux"], P["aux_done"]))
        self.net_forward(coords, n_dev)
        if P["at"] == "fwd":
            P["mid"].record(main)
        rgb, loss, _ = ops.composite_loss_bwd(self.net_out, coords, F["numsteps"], F["ns_c"], F["bg"], F["target"], s.density_grid_mean,
                                              delta=self.loss_func.delta, cascades=s.NERF_CASCADES, dnet=self.dnet,
                                              reg_scale=float(self.world_size))    # the exchange applies 1 / W to the summed gradients
        if zeroing:
            main.wait_event(P["aux_done"])
        self.net_backward(coords, n_dev)
        if P["at"] == "bwd":
            P["mid"].record(main)
        lr = self.optimizer.advance_lr()
        adam = self.optimizer._nested_optimizer
        adam.n_step += 1
        self.ema_optimizer.steps += 1
        # fused Adam+EMA sweeps (optims/adam.py + ema.py; runner.py:75-76): the two MLP weight vectors are single-CTA launches that
        # cost a launch latency each -- on a third stream they run beside the table's sweep instead of after it
        m = self.model
        hyper = (lr, adam.n_step, adam.betas[0], adam.betas[1], adam.eps, self.ema_optimizer.decay)
        if self.world_size > 1:
            self._optimizer_step(lr, adam.n_step)                    # gradient exchange fused with the sliced sweep (section 6)
            return self._pipe_finish(F, loss, rgb, batch, next_batch)
        P["bwd_done"].record(main)
        P["aux"].wait_event(P["bwd_done"])
        with torch.cuda.stream(P["aux"]):
            for p_, g_ in ((m.density_mlp.con_weights, self.dwd), (m.rgb_mlp.con_weights, self.dwr)):
                st = self._st[id(p_)]
                ops.adam_ema(p_.data, g_, st.m, st.v, st.master, *hyper, grad_scale=1.0, zero_grad=True)
            P["aux_done"].record(P["aux"])
        st = self._st[id(m.pos_encoder.m_grid)]
        ops.adam_ema(m.pos_encoder.m_grid.data, self.grid_grad, st.m, st.v, st.master, *hyper, grad_scale=1.0, zero_grad=True)
        main.wait_event(P["aux_done"])
        return self._pipe_finish(F, loss, rgb, batch, next_batch)

    def _pipe_finish(self, F, loss, rgb, batch, next_batch):
        cfg, s, P = self.cfg, self.sampler, self._pipe
        i = cfg.m_training_step
        P["back_done"][F["slot"]].record(P["main"])
        self.last_loss, self.last_rgb = loss, rgb
        cfg.m_training_step = i + 1
        self._pipe_last = F                                          # keeps the front's tensors alive until the next step replaces them
        if i % s.update_den_freq == s.update_den_freq - 1:
            self._adapt_ray_batch(F)                                 # after this step's march, before the next front, as in sample()
        # the front of step i+1, unless that step opens with an occupancy-grid update (needs the sweep above) or the caller feeds
        # batches and has not said which one comes next
        if (i + 1) % s.update_den_freq != 0 and (batch is None or next_batch is not None):
            P["pending"] = self._front(i + 1, next_batch if batch is not None else None, prefetch=True)
        return loss

    def train_step(self, batch=None, next_batch=None):
        if self._pipe is not None:
            return self._train_step_pipe(batch, next_batch)
        if batch is None and self._dev_state is not None:
            return self._train_step_dev()
        cfg, s, m = self.cfg, self.sampler, self.model
        i = cfg.m_training_step
        if batch is None:
            ds = self.dataset["train"]
            R = s.n_rays_per_batch                                   # per-rank rays; the global batch is world_size x R (weak scaling)
            pix = ds.next_pixels(R * self.world_size)
            bg = torch.rand((R * self.world_size, 3), device="cuda", generator=self._bg_gen)    # runner.py:66 (global batch, then this rank's rows)
            if self.world_size > 1:
                lo, hi = dp.shard_range(R, self.rank)
                pix, bg = pix[lo:hi], bg[lo:hi]
            img_ids, rays_o, rays_d, target = ops.prepare_batch(pix.contiguous(), ds.W, ds.H, ds.transforms_gpu, ds.focal_lengths, ds.principal,
                                                                ds.image_data, bg.contiguous())          # dataset.py:172-188 + runner.py:68
        else:
            img_ids, rays_o, rays_d, rgba = batch
            R = rays_o.shape[0]
            bg = torch.rand((R * self.world_size, 3), device="cuda", generator=self._bg_gen)
            if self.world_size > 1:
                bg = bg[self.rank * R:(self.rank + 1) * R].contiguous()
            target = ops.blend_target(rgba.contiguous(), bg.contiguous())                  # runner.py:68
        if i % s.update_den_freq == 0:
            self._table_ready()                                      # the occupancy-grid update evaluates the density network
        s.sample(img_ids, rays_o, rays_d, is_training=True, ray_index_offset=dp.shard_range(R, self.rank)[0])  # grid update /16, march, bookkeeping
        self._table_ready()
        coords, n_dev = s.coords_compacted, s.n_samples_dev
        self.net_forward(coords, n_dev)
        rgb, loss, _ = ops.composite_loss_bwd(self.net_out, coords, s._rays_numsteps, s._rays_numsteps_compacted, bg, target.contiguous(),
                                              s.density_grid_mean, delta=self.loss_func.delta, cascades=s.NERF_CASCADES, dnet=self.dnet,
                                              reg_scale=float(self.world_size))    # the exchange applies 1 / W to the summed gradients
        self.net_backward(coords, n_dev)
        # local loss_scale is 128/R_local (calc_rgb.h:100-101): the all-reduced sum is W x the global-batch gradient
        lr = self.optimizer.advance_lr()
        adam = self.optimizer._nested_optimizer
        adam.n_step += 1
        self.ema_optimizer.steps += 1
        self._optimizer_step(lr, adam.n_step)
        self.last_loss, self.last_rgb = loss, rgb
        cfg.m_training_step = i + 1
        return loss

    # ------------------------------------------------------------------------------------------ host-fed batches
    def train_step_host(self, batch, next_batch=None):
        """One training step on a ray batch that lives in PINNED HOST memory -- (img_ids int32 (R,), rays_o (R,3), rays_d (R,3),
        rgba (R,4) f32), what the reference's dataset yields (dataset.py:172-188).  The batch is copied on a side stream into one of
        two device staging slots; passing `next_batch` starts the copy of the following step's batch before this step's kernels are
        enqueued, so that the copy runs under them.  The mean loss goes back to the host on the side stream as well: the returned
        pinned 1-element tensor holds it once that stream has caught up (read it a step later, or synchronise)."""
        st = self._host_stage
        if st is None:
            st = self._host_stage = dict(stream=torch.cuda.Stream(), slots=[None, None], ready=[torch.cuda.Event(), torch.cuda.Event()],
                                         free=[torch.cuda.Event(), torch.cuda.Event()], staged=[None, None], dev=[None, None], k=0,
                                         loss_host=torch.zeros(1, dtype=torch.float32).pin_memory(), done=torch.cuda.Event())
            for e in st["free"]:
                e.record()
        main = torch.cuda.current_stream()

        def put(b, slot):
            R = b[1].shape[0]
            if st["slots"][slot] is None or st["slots"][slot][1].shape[0] < R:
                cap = max(2 * R, 4096)
                st["slots"][slot] = (torch.empty(cap, dtype=torch.int32, device="cuda"), torch.empty((cap, 3), device="cuda"),
                                     torch.empty((cap, 3), device="cuda"), torch.empty((cap, 4), device="cuda"))
            with torch.cuda.stream(st["stream"]):
                st["stream"].wait_event(st["free"][slot])              # the step that read this slot has finished with it
                for dst, src in zip(st["slots"][slot], b):
                    dst[:R].copy_(src, non_blocking=True)
                st["ready"][slot].record()
            st["staged"][slot] = (b, R)
            st["dev"][slot] = tuple(t[:R] for t in st["slots"][slot])

        slot = st["k"] % 2
        if st["staged"][slot] is None or st["staged"][slot][0] is not batch:
            put(batch, slot)
        if next_batch is not None:
            put(next_batch, 1 - slot)
        main.wait_event(st["ready"][slot])
        dev, nxt = st["dev"][slot], None
        if next_batch is not None and self._pipe is not None:
            nxt = st["dev"][1 - slot]                                # its front (blend, march) is enqueued under this step's kernels
            self._pipe["stream"].wait_event(st["ready"][1 - slot])
        loss = self.train_step(dev, nxt)
        st["free"][slot].record(main)
        st["staged"][slot] = None
        st["k"] += 1
        st["done"].record(main)
        with torch.cuda.stream(st["stream"]):
            st["stream"].wait_event(st["done"])
            loss.record_stream(st["stream"])
            st["loss_host"].copy_(loss.mean().reshape(1), non_blocking=True)
        return st["loss_host"]

    def net_forward(self, coords, n_dev):
        """Fused hash encode + SH + both MLPs on the sampler's coordinate rows -> self.net_out (+ self.enc)."""
        m = self.model
        ops.network_fwd(coords, m.pos_encoder.m_grid, m.pos_encoder.levels, m.density_mlp.con_weights, m.rgb_mlp.con_weights,
                        n_dev=n_dev, out=self.net_out, enc=self.enc)

    def net_backward(self, coords, n_dev):
        """self.dnet -> gradients of the hash table (self.grid_grad) and of both weight vectors (self.dwd, self.dwr)."""
        m = self.model
        ops.network_bwd(coords, self.enc, m.pos_encoder.levels, m.density_mlp.con_weights, m.rgb_mlp.con_weights, self.dnet,
                        self.grid_grad, self.dwd, self.dwr, n_dev=n_dev)

    def _optimizer_step(self, lr, n_step):
        """Gradient exchange + fused Adam/EMA sweep(s) (optims/adam.py + ema.py; runner.py:75-76)."""
        m = self.model
        adam = self.optimizer._nested_optimizer
        hyper = (lr, n_step, adam.betas[0], adam.betas[1], adam.eps, self.ema_optimizer.decay)
        if self.world_size > 1 and self.dp_mode == "p2p":
            st = self._st[id(m.pos_encoder.m_grid)]
            self._epoch += 1
            ops.dp_exchange_step(self.world_size, self.rank, self._hi - self._lo, self._w_param.numel(), self._peers["table"],
                                 self._peers["table_grad"], self._peers["w_grad"], self._peers["flags"], self._epoch, st.m, st.v, st.master,
                                 self._w_param, self._w_m, self._w_v, self._w_master, *hyper, grad_scale=1.0 / self.world_size)
            self._pending_epoch = self._epoch
            return
        if self.world_size > 1:
            # hash table: reduce-scatter -> Adam+EMA on this rank's slice -> async all-gather (waited for in the NEXT step, after the march)
            dp.reduce_scatter_sum(self._slice_grad, self._table_grad, self.pg, self.world_size, self.rank)
            scale = dp.allreduce_grads((self.w_grad,), self.pg, self.world_size)
            st = self._st[id(m.pos_encoder.m_grid)]
            ops.adam_ema(self._slice_param, self._slice_grad, st.m, st.v, st.master, *hyper, grad_scale=scale, zero_grad=False)
            self._table_work = dp.all_gather_slices(self._table, self._slice_param, self.pg, self.world_size, async_op=True)
            self._table_grad.zero_()
            tensors = ((m.density_mlp.con_weights, self.dwd), (m.rgb_mlp.con_weights, self.dwr))
        else:
            scale = 1.0
            tensors = ((m.pos_encoder.m_grid, self.grid_grad), (m.density_mlp.con_weights, self.dwd), (m.rgb_mlp.con_weights, self.dwr))
        for p, g in tensors:
            st = self._st[id(p)]
            ops.adam_ema(p.data, g, st.m, st.v, st.master, *hyper, grad_scale=scale, zero_grad=True)

    # --------------------------------------------------------------------- per-operator path (as the reference wires it)
    def train_step_autograd(self, batch=None):
        cfg = self.cfg
        i = cfg.m_training_step
        img_ids, rays_o, rays_d, rgba = self.next_batch() if batch is None else batch
        bg = torch.rand((rays_o.shape[0], 3), device="cuda", generator=self._bg_gen)
        target = (rgba[:, :3] * rgba[:, 3:] + bg * (1 - rgba[:, 3:])).detach()
        pos, dir_ = self.sampler.sample(img_ids, rays_o, rays_d, is_training=True)
        network_outputs = self.model(pos, dir_)
        rgb = self.sampler.rays2rgb(network_outputs, bg)
        loss = self.loss_func(rgb, target)
        self.optimizer.step(loss)
        self.ema_optimizer.ema_step()
        cfg.m_training_step = i + 1
        return loss.sum(-1)

    def train(self, steps=None, log_every=0):
        end = self.tot_train_steps if steps is None else self.cfg.m_training_step + steps
        step_fn = self.train_step if self.fast else self.train_step_autograd
        while self.cfg.m_training_step < end:
            loss = step_fn()
            i = self.cfg.m_training_step
            if log_every and i % log_every == 0 and self.rank == 0:
                print(f"STEP={i} | LOSS={loss.mean().item():.6f} | rays/batch={self.sampler.n_rays_per_batch}", flush=True)

    # ------------------------------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def render_img(self, dataset_mode="train", img_id=0):
        """runner.py:197-236: tile the image in n_rays_per_batch chunks; returns (img HxWx3, target HxWx3)."""
        self._table_ready()
        self._sync_front()
        ds = self.dataset[dataset_mode]
        W, H = ds.resolution
        rays_o, rays_d = ds.generate_rays_total_test(img_id)
        tile = self.cfg.n_rays_per_batch
        img = torch.empty((H * W, 3), device="cuda")
        alpha = torch.empty((H * W, 1), device="cuda")
        ids = torch.zeros(tile, dtype=torch.int32, device="cuda")
        s, m = self.sampler, self.model
        for p in range(0, H * W, tile):
            e = min(p + tile, H * W)
            o, d = rays_o[p:e], rays_d[p:e]
            if e - p < tile:
                o = torch.cat([o, torch.ones((tile - (e - p), 3), device="cuda")])
                d = torch.cat([d, torch.ones((tile - (e - p), 3), device="cuda")])
            s.sample(ids, o.contiguous(), d.contiguous())
            coords = s._coords
            if self.fast:
                out, _ = ops.network_fwd(coords.contiguous(), m.pos_encoder.m_grid, m.pos_encoder.levels, m.density_mlp.con_weights,
                                         m.rgb_mlp.con_weights, save_enc=False) if coords.shape[0] else (torch.empty((0, 4), dtype=torch.float16, device="cuda"), None)
            else:
                out = m(coords[:, :3], coords[:, 4:])
            rgb, a = s.rays2rgb(out, inference=True)
            img[p:e], alpha[p:e] = rgb[:e - p], a[:e - p]
        bgc = torch.tensor(self.background_color, dtype=torch.float32, device="cuda")
        img = img + bgc * (1 - alpha)
        tar = ds.rgba_for(torch.arange(H * W, device="cuda", dtype=torch.int32) + int(img_id) * H * W)
        tar = tar[:, :3] * tar[:, 3:] + bgc * (1 - tar[:, 3:])
        return img.reshape(H, W, 3), tar.reshape(H, W, 3)

    @torch.no_grad()
    def render_img_nosync(self, dataset_mode="train", img_id=0):
        """N3: the image of render_img without the reference tiler's host round trips (runner.py:206-228 reads the
        sample count back with .item() and copies every tile to the host; 157 tiles per 800x800 image).  The march's device-side
        counter bounds the fused network kernel (n_dev), ngp_composite_infer writes into the image buffer, nothing is read back
        before the caller uses the result.  Same kernels, same RNG consumption, same pixels as render_img."""
        assert self.fast, "render_img_nosync drives the fused network kernel"
        self._table_ready()
        self._sync_front()
        ds = self.dataset[dataset_mode]
        W, H = ds.resolution
        rays_o, rays_d = ds.generate_rays_total_test(img_id)
        tile = self.cfg.n_rays_per_batch
        n_pix = H * W
        n_pad = (n_pix + tile - 1) // tile * tile
        if n_pad > n_pix:                                              # the reference pads the last tile with ones (runner.py:213-217)
            fill = torch.ones((n_pad - n_pix, 3), device="cuda")
            rays_o, rays_d = torch.cat([rays_o, fill]), torch.cat([rays_d, fill])
        img = torch.empty((n_pad, 3), device="cuda")
        alpha = torch.empty((n_pad, 1), device="cuda")
        s, m = self.sampler, self.model
        if getattr(self, "_infer_net_out", None) is None:
            self._infer_net_out = torch.empty((s.max_samples, 4), dtype=torch.float16, device="cuda")
        s._ensure_march_ws(tile)
        for p in range(0, n_pad, tile):
            coords, _, numsteps, counters = ops.march(
                rays_o[p:p + tile].contiguous(), rays_d[p:p + tile].contiguous(), s.density_grid_bitfield, s.aabb_range, s.max_samples,
                s.cone_angle_constant, s.near_distance, s.NERF_CASCADES, s.const_dt, s.rng, coords=s._coords_raw, workspace=s._march_ws)
            ops.pcg32_advance(s.rng)                                   # rng.advance(), ray_sampler.py:61
            ops.network_fwd(coords, m.pos_encoder.m_grid, m.pos_encoder.levels, m.density_mlp.con_weights, m.rgb_mlp.con_weights,
                            n_dev=counters[1:2], save_enc=False, out=self._infer_net_out)
            rgb, a = ops.composite_infer(self._infer_net_out, coords, numsteps, s.NERF_CASCADES)
            img[p:p + tile], alpha[p:p + tile] = rgb, a
        bgc = torch.tensor(self.background_color, dtype=torch.float32, device="cuda")
        img = img[:n_pix] + bgc * (1 - alpha[:n_pix])
        tar = ds.rgba_for(torch.arange(n_pix, device="cuda", dtype=torch.int32) + int(img_id) * n_pix)
        tar = tar[:, :3] * tar[:, 3:] + bgc * (1 - tar[:, 3:])
        return img.reshape(H, W, 3), tar.reshape(H, W, 3)

    @torch.no_grad()
    def psnr(self, dataset_mode="val", max_images=None):
        """mean over images of -10 log10(mse) (runner.py:86-99, mse_loss.py:6-7)."""
        ds = self.dataset[dataset_mode]
        n = ds.n_images if max_images is None else min(max_images, ds.n_images)
        tot = 0.0
        for k in range(n):
            img, tar = self.render_img(dataset_mode, k)
            tot += float(L.mse2psnr(L.img2mse(img, tar)).item())
        return tot / n

    # ------------------------------------------------------------------------------------------ checkpoint (N4)
    def save_ckpt(self, path):
        """Every rank must call this when world_size > 1 (the table's optimizer state is gathered); rank 0 writes the file."""
        adam = self.optimizer._nested_optimizer
        self._table_ready()
        nested = adam.state_dict()
        if self.world_size > 1:
            k = [id(s.p) for s in adam.state].index(id(self.model.pos_encoder.m_grid))
            full = self._gather_table_state()
            for key, t in zip(("m", "v", "master"), full):
                nested[key] = list(nested[key])
                nested[key][k] = t
            if self.rank != 0:
                return
        sampler_state = self.sampler.state_dict()
        if self._pipe is not None and self._pipe["pending"] is not None:
            # a prefetched front has drawn the next step's jitter already: the checkpoint holds the stream position of global_step
            sampler_state["rng"] = torch.from_numpy(self._pipe["pending"]["rng_before"].astype(np.int64))
        ck = {"global_step": self.cfg.m_training_step, "model": self.model.state_dict(), "sampler": sampler_state,
              "optimizer": self.optimizer.state_dict(), "nested_optimizer": nested, "ema_optimizer": self.ema_optimizer.state_dict()}
        if str(path).endswith(".pkl") and not hasattr(self.model.density_mlp, "con_weights"):
            raise NotImplementedError("the .pkl interchange format is written for the fused-MLP parameter layout (con_weights); "
                                      "save the nn.Linear fallback model to a .pt path")
        if str(path).endswith(".pkl"):
            # the reference's params.pkl wire format (runner/runner.py:123-131), readable by its load_ckpt (:133-151)
            from .utils import ckpt_compat as cc
            dec = self.optimizer
            ref = cc.native_to_reference(
                ck, adam_hyper=dict(lr=adam.lr, eps=adam.eps, betas=tuple(adam.betas)),
                expdecay_hyper=dict(base_lr=dec.base_lr, decay_start=dec.decay_start, decay_interval=dec.decay_interval, decay_base=dec.decay_base,
                                    decay_end=dec.decay_end),
                param_dtype=np.float16 if self.model.pos_encoder.m_grid.dtype == torch.float16 else np.float32)
            cc.write_reference_ckpt(ref, path)
            return
        torch.save(ck, path)

    def load_ckpt(self, path):
        self._table_ready()                                          # an exchange of the previous step may still be writing the table
        if self._pipe is not None:
            self._sync_front()
            self._pipe["pending"] = None                             # marched against the occupancy grid that is about to be replaced
        if self.world_size > 1:
            import torch.distributed as dist
            dist.barrier(group=self.pg)                              # no peer may still push into this rank's table while it is overwritten
        if str(path).endswith(".pkl") and not hasattr(self.model.density_mlp, "con_weights"):
            raise NotImplementedError("the .pkl interchange format carries the fused-MLP parameter layout; load a .pt checkpoint instead")
        if str(path).endswith(".pkl"):
            from .utils import ckpt_compat as cc
            ck = cc.load_native_from_reference_file(path, self.model.pos_encoder.m_grid.numel(), "cuda",
                                                    {k: v.dtype for k, v in self.model.state_dict().items()})
        else:
            ck = torch.load(path, map_location="cuda", weights_only=True)      # tensors, numbers, lists and dicts only
        self.cfg.m_training_step = self.start = ck["global_step"]
        self.model.load_state_dict(ck["model"])
        self.sampler.load_state_dict(ck["sampler"])
        self.optimizer.load_state_dict(ck["optimizer"])
        nested = ck["nested_optimizer"]
        if self.world_size > 1:
            adam = self.optimizer._nested_optimizer
            k = [id(s.p) for s in adam.state].index(id(self.model.pos_encoder.m_grid))
            self._scatter_table_state(nested["m"][k], nested["v"][k], nested["master"][k])
            st = self._st[id(self.model.pos_encoder.m_grid)]
            nested = dict(nested, **{key: [t if j != k else getattr(st, key) for j, t in enumerate(nested[key])] for key in ("m", "v", "master")})
            self._slice_param.copy_(self._table[self._lo:self._hi])
        self.optimizer._nested_optimizer.load_state_dict(nested)
        self.ema_optimizer.load_state_dict(ck["ema_optimizer"])


def lego_cfg(fp16=True, synthetic=True, **over):
    """projects/ngp/configs/ngp_base.py key for key, + fp16 (BASELINE config #2) and the synthetic stand-in dataset."""
    ds_type = "SyntheticNerfDataset" if synthetic else "NerfDataset"
    c = dict(
        sampler=dict(type="DensityGridSampler", update_den_freq=16),
        encoder=dict(pos_encoder=dict(type="HashEncoder"), dir_encoder=dict(type="SHEncoder")),
        model=dict(type="NGPNetworks", use_fully=True),
        loss=dict(type="HuberLoss", delta=0.1),
        optim=dict(type="Adam", lr=1e-1, eps=1e-15, betas=(0.9, 0.99)),
        ema=dict(type="EMA", decay=0.95),
        expdecay=dict(type="ExpDecay", decay_start=20_000, decay_interval=10_000, decay_base=0.33, decay_end=None),
        dataset=dict(train=dict(type=ds_type, root_dir="data/lego", batch_size=4096, mode="train"),
                     val=dict(type=ds_type, root_dir="data/lego", batch_size=4096, mode="val", preload_shuffle=False),
                     test=dict(type=ds_type, root_dir="data/lego", batch_size=4096, mode="test", preload_shuffle=False)),
        exp_name="lego", log_dir="./logs", tot_train_steps=40000, background_color=[0, 0, 0],
        hash_func="p0 ^ p1 * 19349663 ^ p2 * 83492791", cone_angle_constant=0.00390625, near_distance=0.2, n_rays_per_batch=4096,
        n_training_steps=16, target_batch_size=1 << 18, const_dt=True, load_ckpt=False, ckpt_path=None, alpha_image=False, fp16=fp16,
    )
    c.update(over)
    return c


def fox_cfg(fp16=True, synthetic=True, **over):
    """projects/ngp/configs/ngp_fox.py key for key (BASELINE config #3: aabb_scale 4 from the capture's transforms, cone stepping
    `const_dt=False`, fp16, no validation split); `synthetic` swaps data/fox for the procedural stand-in with the capture's
    resolution, intrinsics, camera arc and aabb_scale (plugin/dataset.py: SyntheticNerfDataset(style="fox"))."""
    c = lego_cfg(fp16=fp16, synthetic=synthetic)
    ds_type = "SyntheticNerfDataset" if synthetic else "NerfDataset"
    extra = dict(style="fox") if synthetic else {}
    c.update(dataset=dict(train=dict(type=ds_type, root_dir="data/fox", batch_size=4096, mode="train", **extra),
                          test=dict(type=ds_type, root_dir="data/fox", batch_size=4096, mode="test", preload_shuffle=False, **extra)),
             exp_name="fox", const_dt=False)
    c.update(over)
    return c
