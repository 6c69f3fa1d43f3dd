"""Drop-in surface for the hot path: the reference's call signatures over the B200 engines.

  * `B200BrushNetModel.forward`      <- BrushNetModel.forward          (S/models/brushnet.py:678-692)
  * `B200UNet2DConditionModel.forward` <- UNet2DConditionModel.forward (S/models/unets/unet_2d_condition.py:1039-1057)
  * `MirrorFusionB200Pipeline.__call__` <- the denoise loop of StableDiffusionBrushNetPipeline.__call__
                                         (S/pipelines/brushnet/pipeline_brushnet.py:848-880, 1219-1332)
  * `StepEngine`: the fused form the pipeline actually runs: BrushNet + UNet + CFG + scheduler step as one
    CUDA graph per geometry, taps handed over in HBM buffers shared between the two engines.

Tensors cross this boundary in the reference's layout (NCHW, any float dtype); inside everything is NHWC bf16.
VAE / CLIP / image preprocessing stay outside: the pipeline takes `prompt_embeds` and
`conditioning_latents` (or a `vae_encode` callable) like the parity recipe of SURVEY.md Appendix A.
"""
from __future__ import annotations

import os

import inspect
from types import SimpleNamespace
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch

from . import ops
from .config import NetConfig, SD15, tap_channels
from .engine import BrushNetEngine, UNetEngine
from .schedulers import B200DDIMScheduler, B200UniPCScheduler

f32 = torch.float32


def _as_timestep_vector(timestep, batch: int, device) -> torch.Tensor:
    """Scalar / 0-dim / [B] timestep -> fp32 [B] on device (unet_2d_condition.py:1141-1152)."""
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([float(timestep)], dtype=f32)
    t = timestep.to(device=device, dtype=f32).reshape(-1)
    if t.numel() == 1:
        t = t.expand(batch)
    if t.numel() != batch:
        raise ValueError(f"timestep has {t.numel()} entries for a batch of {batch}")
    return t.contiguous()


class _GeomCache:
    def __init__(self, factory):
        self._factory = factory
        self._engines: Dict[Tuple[int, int, int], Any] = {}

    def get(self, B, H, W):
        key = (B, H, W)
        if key not in self._engines:
            self._engines[key] = self._factory(B, H, W)
        return self._engines[key]


class B200BrushNetModel:
    """BrushNetModel on sm_100a kernels.  `forward` keeps the reference signature and return convention:
    `(down_block_res_samples: list[12], mid_block_res_sample, up_block_res_samples: list[15])`, fresh lists each call
    (the UNet pops them), NCHW in the dtype of `sample`."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: NetConfig = SD15, device="cuda", dtype=torch.bfloat16):
        self.cfg = cfg
        self.device = torch.device(device)
        self.dtype = dtype
        self.config = SimpleNamespace(global_pool_conditions=False, in_channels=cfg.in_channels,
                                      conditioning_channels=cfg.conditioning_channels,
                                      block_out_channels=cfg.block_out_channels)
        self._sd = state_dict
        self._cache = _GeomCache(lambda B, H, W: BrushNetEngine(cfg, state_dict, B, H, W, self.device))

    def engine(self, B, H, W) -> BrushNetEngine:
        return self._cache.get(B, H, W)

    def forward(self, sample, timestep, encoder_hidden_states=None, brushnet_cond=None, conditioning_scale: float = 1.0,
                class_labels=None, timestep_cond=None, attention_mask=None, added_cond_kwargs=None,
                cross_attention_kwargs=None, guess_mode: bool = False, return_dict: bool = True):
        if brushnet_cond is None:
            raise ValueError("brushnet_cond is required")
        if class_labels is not None or timestep_cond is not None or added_cond_kwargs:
            raise NotImplementedError("class / additional embeddings are not on the MirrorFusion path")
        B, _, H, W = sample.shape
        if brushnet_cond.shape[1] != self.cfg.conditioning_channels:
            raise ValueError(f"brushnet_cond has {brushnet_cond.shape[1]} channels, expected {self.cfg.conditioning_channels}")
        e = self.engine(B, H, W)
        e.sample_in.copy_(sample)
        e.cond_in.copy_(brushnet_cond)
        e.t_dev.copy_(_as_timestep_vector(timestep, B, self.device))
        e.set_scale(float(conditioning_scale), guess_mode=bool(guess_mode))     # guess mode: log-spaced tap scales (brushnet.py:896-902)
        e.run()
        outs = []
        for t, (h, w) in zip(e.taps, e.tap_hw):
            o = torch.empty(B, t.shape[-1], h, w, device=self.device, dtype=f32)
            ops.nhwc_to_nchw(t, o)
            outs.append(o.to(sample.dtype))
        down, mid, up = outs[: e.n_down], outs[e.n_down], outs[e.n_down + 1:]
        if not return_dict:
            return down, mid, up
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid, up_block_res_samples=up)

    __call__ = forward


class B200UNet2DConditionModel:
    """UNet2DConditionModel (SD1.5 family) on sm_100a kernels with the three BrushNet kwargs of the fork
    (unet_2d_condition.py:1054-1056).  Tap lists are consumed with pop(0) like the reference (:1218,1228,1306)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: NetConfig = SD15, device="cuda", dtype=torch.bfloat16):
        self.cfg = cfg
        self.device = torch.device(device)
        self.dtype = dtype
        self.config = SimpleNamespace(in_channels=cfg.in_channels, time_cond_proj_dim=None, sample_size=cfg.sample_size,
                                      block_out_channels=cfg.block_out_channels)
        self._sd = state_dict
        self._cache = _GeomCache(lambda B, H, W: UNetEngine(cfg, state_dict, B, H, W, self.device))
        self._ctx_key = None

    def engine(self, B, H, W) -> UNetEngine:
        return self._cache.get(B, H, W)

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs=None, added_cond_kwargs=None, down_block_additional_residuals=None,
                mid_block_additional_residual=None, down_intrablock_additional_residuals=None,
                encoder_attention_mask=None, return_dict: bool = True, down_block_add_samples=None,
                mid_block_add_sample=None, up_block_add_samples=None):
        if any(a is not None for a in (class_labels, timestep_cond, attention_mask, added_cond_kwargs,
                                       down_block_additional_residuals, mid_block_additional_residual,
                                       down_intrablock_additional_residuals, encoder_attention_mask)):
            raise NotImplementedError("ControlNet / adapter / mask arguments are not on the MirrorFusion path")
        B, _, H, W = sample.shape
        e = self.engine(B, H, W)
        key = (encoder_hidden_states.data_ptr(), encoder_hidden_states._version, tuple(encoder_hidden_states.shape), id(e))
        if key != self._ctx_key:
            if encoder_hidden_states.shape[1] != e.ctx_len:
                raise ValueError(f"encoder_hidden_states must have {e.ctx_len} tokens")
            e.set_context(encoder_hidden_states)
            self._ctx_key = key
        is_brushnet = down_block_add_samples is not None and mid_block_add_sample is not None \
            and up_block_add_samples is not None                                          # :1202
        if is_brushnet:
            n_up = len(e.taps) - e.n_down - 1
            if len(down_block_add_samples) != e.n_down or len(up_block_add_samples) != n_up:
                raise ValueError("wrong number of BrushNet residuals")
            srcs = [down_block_add_samples.pop(0) for _ in range(e.n_down)] + [mid_block_add_sample] + \
                   [up_block_add_samples.pop(0) for _ in range(n_up)]
            for dst, s in zip(e.taps, srcs):
                ops.nchw_to_nhwc(s.to(f32).contiguous(), dst)
        else:
            for dst in e.taps:
                dst.zero_()
        e.sample_in.copy_(sample)
        e.t_dev.copy_(_as_timestep_vector(timestep, B, self.device))
        e.run()
        out = e.out.to(sample.dtype).clone()
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)

    __call__ = forward


class B200AttnProcessor:
    """Attention processor on the sm_100a kernels with the reference's processor call contract
    `proc(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0)`
    (AttnProcessor2_0.__call__, S/models/attention_processor.py:1213-1286), installable with
    `Attention.set_processor` / `unet.set_attn_processor` (:380-398, unet_2d_condition.py:716-748).
    It reads `attn.to_q/to_k/to_v/to_out[0]`, `attn.heads`, `attn.residual_connection`, `attn.rescale_output_factor`;
    q/k/v run as ONE fused GEMM for self-attention, the softmax(QK^T/sqrt(d))V is the flash tcgen05 kernel.
    Plans and packed weights are cached per (module, geometry)."""

    def __init__(self):
        self._cache: Dict[Tuple, Any] = {}

    def _state(self, attn, B, T, Tk, C, Cctx, is_self, dev):
        key = (id(attn), B, T, Tk, C, Cctx, is_self)
        st = self._cache.get(key)
        if st is not None:
            return st
        bf = torch.bfloat16
        heads = attn.heads
        z = lambda *shape: torch.zeros(*shape, device=dev, dtype=bf)
        st = SimpleNamespace()
        st.x = z(B * T, C)
        st.att = z(B * T, C)
        st.out = z(B * T, C)
        wq, wk, wv = (attn.to_q.weight.detach(), attn.to_k.weight.detach(), attn.to_v.weight.detach())
        wo = attn.to_out[0].weight.detach().to(device=dev, dtype=bf).contiguous()
        bo = attn.to_out[0].bias
        st.bo = None if bo is None else bo.detach().to(device=dev, dtype=f32).contiguous()
        if is_self:
            st.qkv = z(B * T, 3 * C)
            w = torch.cat([wq, wk, wv], 0).to(device=dev, dtype=bf).contiguous()
            st.plans = [ops.linear_plan(st.x, w, st.qkv)]
        else:
            st.ctx = z(B * Tk, Cctx)
            st.q = z(B * T, C)
            st.k = z(B * Tk, C)
            st.v = z(B * Tk, C)
            st.plans = [ops.linear_plan(st.x, wq.to(device=dev, dtype=bf).contiguous(), st.q),
                        ops.linear_plan(st.ctx, wk.to(device=dev, dtype=bf).contiguous(), st.k),
                        ops.linear_plan(st.ctx, wv.to(device=dev, dtype=bf).contiguous(), st.v)]
        st.out_plan = ops.linear_plan(st.att, wo, st.out, bias=st.bo)
        st.heads = heads
        self._cache[key] = st
        return st

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale: float = 1.0):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not on the MirrorFusion path (the pipeline never passes one)")
        if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "group_norm", None) is not None \
                or getattr(attn, "norm_cross", None):
            raise NotImplementedError("spatial_norm / group_norm / norm_cross attention variants are not on the MirrorFusion path")
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b_, c_, h_, w_ = hidden_states.shape
            hidden_states = hidden_states.view(b_, c_, h_ * w_).transpose(1, 2)
        B, T, C = hidden_states.shape
        is_self = encoder_hidden_states is None
        Tk = T if is_self else encoder_hidden_states.shape[1]
        Cctx = C if is_self else encoder_hidden_states.shape[2]
        heads = attn.heads
        d = C // heads
        dev = hidden_states.device
        st = self._state(attn, B, T, Tk, C, Cctx, is_self, dev)
        st.x.copy_(hidden_states.reshape(B * T, C))
        if is_self:
            st.plans[0].run()
            ops.attention(st.qkv, st.qkv.view(-1)[C:], st.qkv.view(-1)[2 * C:], st.att, B=B, heads=heads, head_dim=d, Tq=T, Tk=T,
                          ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C)
        else:
            st.ctx.copy_(encoder_hidden_states.reshape(B * Tk, Cctx))
            for p in st.plans:
                p.run()
            ops.attention(st.q, st.k, st.v, st.att, B=B, heads=heads, head_dim=d, Tq=T, Tk=Tk, ldq=C, ldk=C, ldv=C, ldo=C)
        st.out_plan.run()
        out = st.out.view(B, T, C).to(hidden_states.dtype)
        if input_ndim == 4:
            out = out.transpose(-1, -2).reshape(b_, c_, h_, w_)
        if getattr(attn, "residual_connection", False):
            out = out + residual
        return out / getattr(attn, "rescale_output_factor", 1.0)


class StepEngine:
    """One fused denoise step for `images` images (net batch 2*images, [uncond, cond] halves):
    latents -> BrushNet -> UNet (+taps) -> CFG -> scheduler, all state resident on the device."""

    def __init__(self, cfg: NetConfig, unet_sd, brushnet_sd, images: int, H: int, W: int, device="cuda",
                 use_graph: bool = True, fuse_taps: bool = True, two_streams: bool = False,
                 dedup_brushnet_cfg: bool = False, precision: str = "bf16", host_pack: bool = False, guess_mode: bool = False):
        """precision: "bf16" = the product path (tcgen05 kernels); "fp32" = the PARITY MODE of BASELINE config 1 — the
        same program (fusions, K-segments, tap folding, buffers) with fp32 storage on the CUDA-core kernels of
        csrc/fp32mode.cu, for the rel-L2 1e-4 bar against the fp32 reference.
        dedup_brushnet_cfg (opt-in, exact): BrushNetModel has no attention, so `encoder_hidden_states` is unused
        (brushnet.py:678-925) and in the pipeline's default mode both CFG halves of its batch are identical by
        construction (`latent_model_input = torch.cat([latents] * 2)`, doubled `conditioning_latents`,
        pipeline_brushnet.py:1256,1188-1202).  The branch is then evaluated on `images` samples and its 28 features
        are broadcast to both halves before the UNet consumes them — bit-identical taps, 18 % fewer FLOPs per step.
        `set_conditioning` refuses conditioning whose halves differ.  Off by default: the headline numbers of
        bench.py run the reference's full 2b-sample BrushNet.
        guess_mode: the pipeline's guess mode with CFG (pipeline_brushnet.py:1262-1301): BrushNet runs on the CONDITIONAL batch
        only (`images` samples), its 28 taps are scaled by conditioning_scale x logspace(-1, 0, 28) (brushnet.py:896-902) and the
        unconditional half of the UNet receives zeros — here the zero-convs write straight into the second half of the UNet's tap
        buffers, whose first half stays zero."""
        self.guess_mode = bool(guess_mode)
        if self.guess_mode and (dedup_brushnet_cfg or two_streams):
            raise ValueError("guess_mode excludes dedup_brushnet_cfg / two_streams")
        self.precision = precision
        self._hp = bool(host_pack)          # engine._Net: repack the weights on the host and upload the results
        with ops.precision(precision):      # engines read the storage dtype while they are built
            self._build(cfg, unet_sd, brushnet_sd, images, H, W, device, use_graph, fuse_taps, two_streams, dedup_brushnet_cfg)

    def _build(self, cfg, unet_sd, brushnet_sd, images, H, W, device, use_graph, fuse_taps, two_streams, dedup_brushnet_cfg):
        hp = self._hp
        self.cfg, self.images, self.H, self.W = cfg, images, H, W
        self.dev = torch.device(device)
        B = 2 * images
        self.fuse_taps = fuse_taps
        self.dedup = bool(dedup_brushnet_cfg)
        if self.dedup and not fuse_taps:
            raise ValueError("dedup_brushnet_cfg requires fuse_taps=True")
        if self.guess_mode:
            self.fuse_taps = fuse_taps = False
            self.unet = UNetEngine(cfg, unet_sd, B, H, W, self.dev, host_pack=hp)
            self.bn = BrushNetEngine(cfg, brushnet_sd, images, H, W, self.dev, tap_bufs=[t[images:] for t in self.unet.taps], host_pack=hp)
        elif self.dedup:
            self.bn = BrushNetEngine(cfg, brushnet_sd, images, H, W, self.dev, only_first_tap=True, host_pack=hp, dup_halves=True)
            self.unet = UNetEngine(cfg, unet_sd, B, H, W, self.dev, tap_sources=self.bn.dup_sources, tap0=self.bn.dup_tap0, host_pack=hp)
        elif fuse_taps:
            # 27 of the 28 zero-convs run inside the UNet GEMM that consumes the tap (extra K-segment); only the
            # conv_in-site tap is a tensor
            self.bn = BrushNetEngine(cfg, brushnet_sd, B, H, W, self.dev, only_first_tap=True, host_pack=hp)
            self.unet = UNetEngine(cfg, unet_sd, B, H, W, self.dev, tap_sources=self.bn.tap_sources, tap0=self.bn.taps[0], host_pack=hp)
        else:
            self.unet = UNetEngine(cfg, unet_sd, B, H, W, self.dev, host_pack=hp)
            self.bn = BrushNetEngine(cfg, brushnet_sd, B, H, W, self.dev, tap_bufs=self.unet.taps, host_pack=hp)
        self._tap_scale = 1.0
        self.two_streams = two_streams
        self._side_stream = torch.cuda.Stream(device=self.dev) if two_streams else None
        self.time_tables = None     # (timesteps, brushnet table, unet table) once prepare_timesteps() has run
        z = lambda: torch.zeros(images, cfg.in_channels, H, W, device=self.dev, dtype=f32)
        self.x, self.last, self.m0, self.m1 = z(), z(), z(), z()
        self.coef = torch.zeros(12, device=self.dev, dtype=f32)
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        # eager mode (use_graph=False): replay the step as a program recorded inside the library (ops.Program) instead of walking
        # the Python launch list; MFB_NATIVE_PROGRAM=0 keeps the Python loop (A/B, debugging)
        self.native_program = os.environ.get("MFB_NATIVE_PROGRAM", "1") == "1"
        self._program: Optional[ops.Program] = None
        self.launches_per_step = self.unet.launches + self.bn.launches + 1
        self.flops_per_step = self.unet.flops + self.bn.flops

    def set_conditioning(self, prompt_embeds: torch.Tensor, conditioning_latents: torch.Tensor):
        self.unet.set_context(prompt_embeds)
        if self.guess_mode:        # prepare_image does not duplicate the conditioning in guess mode (:771-772): `images` samples
            n = self.images
            self.bn.cond_in.copy_(conditioning_latents[-n:])
        elif self.dedup:
            n = self.images
            if not torch.equal(conditioning_latents[:n], conditioning_latents[n:]):
                raise ValueError("dedup_brushnet_cfg: the two CFG halves of conditioning_latents differ")
            self.bn.cond_in.copy_(conditioning_latents[:n])
        else:
            self.bn.cond_in.copy_(conditioning_latents)

    def _enqueue(self):
        n = self.images
        for e in (self.bn, self.unet):                    # latent_model_input = cat([latents] * 2) (:1256)
            ops.copy_f32(e.sample_in[:n], self.x)         # (a library kernel: part of the recorded program / graph)
            if e.sample_in.shape[0] > n:                  # (the de-duplicated / guess-mode BrushNet sees each latent once)
                ops.copy_f32(e.sample_in[n:], self.x)
        skip = 0 if self.time_tables is None else None    # hoisted timestep path: row biases were copied in by step()
        if self.two_streams:
            self._run_two_streams(skip)
        else:
            for e in (self.bn, self.unet):
                for f in e.prog[(e.n_time_ops if skip is None else 0):]:
                    f()
        eps = self.unet.out
        ops.cfg_sched_step(eps[:n], eps[n:], self.x, self.last, self.m0, self.m1, self.coef)

    def prepare_timesteps(self, timesteps):
        """Hoist the timestep path (sinusoid -> MLP -> all 22+22 time_emb_proj, embeddings.py:27-67,226-237,
        resnet.py:369-376) out of the loop: it depends only on t, so both nets' row-bias tables for the whole
        schedule are computed once.  The first call must precede the capture of the step graph (the captured program skips
        the timestep ops); later calls — another step count, another scheduler — just rebuild the tables: they are not baked
        into the graph, `step()` copies the row of the current timestep into the row-bias buffers before every replay."""
        if self.graph is not None and self.time_tables is None:
            raise RuntimeError("this step graph was captured with the timestep path inside; it cannot switch to hoisted tables")
        ts = [float(t) for t in timesteps]
        self.time_tables = ({t: i for i, t in enumerate(ts)}, self.bn.timestep_table(ts), self.unet.timestep_table(ts))
        self.launches_per_step = self.unet.launches + self.bn.launches + 1 - 8
        self._program = None          # a recorded program that still holds the timestep ops is re-recorded at the next step

    def _run_two_streams(self, skip):
        """BrushNet on a side stream, UNet on the current one.  The UNet entry that consumes BrushNet tensor k (a fused
        zero-conv K-segment or the conv_in-site tap) waits on an event recorded right after the BrushNet entry that
        writes it, so BrushNet runs ahead and its bandwidth-bound kernels (GroupNorm, boundary convs) overlap the UNet's
        tensor-bound ones and vice versa.  Works eagerly and under CUDA-graph capture (fork/join on the capture stream)."""
        cur = torch.cuda.current_stream()
        side = self._side_stream
        b0 = self.bn.n_time_ops if skip is None else 0
        u0 = self.unet.n_time_ops if skip is None else 0
        # UNet position -> BrushNet position it depends on
        need = {}
        for pos, ptrs in self.unet.ext_reads.items():
            deps = [self.bn.writer_pos[p] for p in ptrs if p in self.bn.writer_pos]
            if deps and pos >= u0:
                need[pos] = max(deps)
        wanted = sorted(set(need.values()))
        events = {bp: torch.cuda.Event() for bp in wanted}
        side.wait_stream(cur)                                        # fork
        with torch.cuda.stream(side):
            for i in range(b0, len(self.bn.prog)):
                self.bn.prog[i]()
                if i in events:
                    events[i].record(side)
        for i in range(u0, len(self.unet.prog)):
            if i in need:
                cur.wait_event(events[need[i]])
            self.unet.prog[i]()
        cur.wait_stream(side)                                        # join

    def step(self, t: float, coef_row: torch.Tensor, scale: float = 1.0):
        if self.time_tables is not None:
            i = self.time_tables[0][float(t)]
            self.bn.rowbias.copy_(self.time_tables[1][i].expand_as(self.bn.rowbias))
            self.unet.rowbias.copy_(self.time_tables[2][i].expand_as(self.unet.rowbias))
        else:
            self.bn.t_dev.fill_(float(t))
            self.unet.t_dev.fill_(float(t))
        self.bn.set_scale(float(scale), guess_mode=self.guess_mode)
        if self.fuse_taps and float(scale) != self._tap_scale:      # rare: only at control-guidance window edges
            self.unet.set_tap_scale(float(scale))
            self._tap_scale = float(scale)
        self.coef.copy_(coef_row, non_blocking=True)
        if not self.use_graph:
            if self.native_program and not (self.two_streams or self.dedup):
                # the whole step from ONE C call: the launch sequence is recorded inside libmfb200 the first time it runs
                # (mfb_program_begin / _end) and replayed by mfb_program_run afterwards — no Python between the 512 launches
                if self._program is None:
                    with ops.Program() as prog:
                        self._enqueue()
                    self._program = prog
                else:
                    self._program.run()
                return
            self._enqueue()
            return
        if self.graph is None:
            self._capture()
        self.graph.replay()

    def _capture(self):
        # one eager pass first (lazy cudaFuncSetAttribute calls, allocator warm-up), on saved state
        saved = [t.clone() for t in (self.x, self.last, self.m0, self.m1)]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._enqueue()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue()
        for dst, src in zip((self.x, self.last, self.m0, self.m1), saved):
            dst.copy_(src)
        self.graph = g

    def denoise(self, latents: torch.Tensor, scheduler, num_inference_steps: int, guidance_scale: float = 7.5,
                conditioning_scales: Optional[List[float]] = None,
                callback: Optional[Callable[[int, int, torch.Tensor], Optional[torch.Tensor]]] = None, eta: float = 0.0,
                generator=None) -> torch.Tensor:
        """eta / generator: the `extra_step_kwargs` of the reference loop (pipeline_brushnet.py:556-571,1315) — used by DDIM only
        (UniPC's `step` takes neither, so the reference drops them).  With eta > 0 every step's variance noise is drawn exactly
        like `randn_tensor` in DDIMScheduler.step does and rides into the fused kernel as its `m0` operand."""
        scheduler.set_timesteps(num_inference_steps, device="cpu")
        stochastic = bool(eta) and getattr(scheduler, "takes_variance_noise", False)
        table = (scheduler.coefficient_table(guidance_scale, eta) if stochastic else scheduler.coefficient_table(guidance_scale)).to(self.dev)
        self.x.copy_(latents.to(device=self.dev, dtype=f32) * scheduler.init_noise_sigma)
        for t_ in (self.last, self.m0, self.m1):
            t_.zero_()
        ts = scheduler.timesteps.tolist()
        if self.graph is None or (self.time_tables is not None and any(float(t) not in self.time_tables[0] for t in ts)):
            self.prepare_timesteps(ts)
        for i, t in enumerate(ts):
            sc = 1.0 if conditioning_scales is None else conditioning_scales[i]
            if stochastic:
                self.m0.copy_(scheduler.variance_noise(self.x.shape, generator, self.dev, f32))
            self.step(float(t), table[i], sc)
            if callback is not None:
                new = callback(i, int(t), self.x)
                if new is not None:
                    self.x.copy_(new)
        return self.x


class MirrorFusionB200Pipeline:
    """The denoise loop of StableDiffusionBrushNetPipeline behind its `__call__` argument names."""

    def __init__(self, unet_state_dict, brushnet_state_dict, scheduler=None, cfg: NetConfig = SD15, device="cuda",
                 depth_conditioning_mode: str = "concat", normals_conditioning_mode: Optional[str] = None,
                 vae_encode: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 vae_decode: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, vae_scale_factor: int = 8,
                 precision: str = "bf16", vae_state_dict=None, vae_cfg=None, empty_prompt_embeds: Optional[torch.Tensor] = None):
        """empty_prompt_embeds: the text encoder's embedding of "" — what the reference substitutes when no negative prompt is given.
        precision="fp32": the fp32 parity mode of StepEngine (BASELINE config 1), otherwise the bf16 product path.
        vae_state_dict (AutoencoderKL.state_dict() keys; only post_quant_conv.* / decoder.* are read): when given and no
        `vae_decode` callable is passed, images are decoded by VaeDecoderEngine on the same kernels (vae.py)."""
        if depth_conditioning_mode != "concat" or normals_conditioning_mode is not None:
            raise NotImplementedError("only depth_conditioning_mode='concat' (the released MirrorFusion checkpoint) is implemented")
        self.cfg, self.device = cfg, torch.device(device)
        self.unet_sd, self.brushnet_sd = unet_state_dict, brushnet_state_dict
        self.scheduler = scheduler or B200UniPCScheduler()
        self.vae_encode, self.vae_decode, self.vae_scale_factor = vae_encode, vae_decode, vae_scale_factor
        self._engines: Dict[Tuple[int, int, int], StepEngine] = {}
        self.precision = precision
        self.empty_prompt_embeds = empty_prompt_embeds
        self.vae_sd, self.vae_cfg, self._vae_engines = vae_state_dict, vae_cfg, {}
        if vae_state_dict is not None and vae_decode is None:
            self.vae_decode = self._decode_on_kernels
        if vae_state_dict is not None and vae_encode is None and any(k.startswith("encoder.") for k in vae_state_dict):
            self.vae_encode = self._encode_on_kernels

    def _encode_on_kernels(self, img: torch.Tensor) -> torch.Tensor:
        """AutoencoderKL.encode(img).latent_dist.sample() (pipeline_brushnet.py:1188-1190) by VaeEncoderEngine."""
        from .vae import SD_VAE, VaeEncoderEngine
        key = ("enc",) + tuple(img.shape)
        if key not in self._vae_engines:
            with ops.precision(self.precision):
                self._vae_engines[key] = VaeEncoderEngine(self.vae_cfg or SD_VAE, self.vae_sd, img.shape[0], img.shape[2], img.shape[3],
                                                          self.device)
        eng = self._vae_engines[key]
        noise = torch.randn(img.shape[0], eng.mean.shape[1], *eng.latent_hw, device=self.device, dtype=torch.float32)
        return eng.encode(img, noise=noise).clone()

    @classmethod
    def from_checkpoint(cls, unet_dir: str, brushnet_dir: str, **kw) -> "MirrorFusionB200Pipeline":
        """Build from the reference's on-disk layout: two diffusers model directories (config.json +
        diffusion_pytorch_model.safetensors), e.g. `<base>/unet` and `checkpoint-N/brushnet`
        (E/train_brushnet_mirror.py:997-1032), read without instantiating torch modules (checkpoint.py)."""
        from .checkpoint import load_mirrorfusion
        cfg, usd, bsd = load_mirrorfusion(unet_dir, brushnet_dir)
        return cls(usd, bsd, cfg=cfg, **kw)

    def _decode_on_kernels(self, z: torch.Tensor) -> torch.Tensor:
        from .vae import SD_VAE, VaeDecoderEngine
        key = tuple(z.shape)
        if key not in self._vae_engines:
            with ops.precision(self.precision):
                self._vae_engines[key] = VaeDecoderEngine(self.vae_cfg or SD_VAE, self.vae_sd, z.shape[0], z.shape[2], z.shape[3],
                                                          self.device)
        return self._vae_engines[key].decode(z).clone()

    def engine(self, images, H, W, guess_mode: bool = False) -> StepEngine:
        key = (images, H, W) + (("guess",) if guess_mode else ())
        if key not in self._engines:
            # BrushNet on a second launch stream (bit-identical, fills the bubbles at kernel transitions): not with guess mode
            self._engines[key] = StepEngine(self.cfg, self.unet_sd, self.brushnet_sd, images, H, W, self.device,
                                            precision=self.precision, guess_mode=guess_mode, two_streams=not guess_mode)
        return self._engines[key]

    def check_inputs(self, prompt_embeds, negative_prompt_embeds, brushnet_conditioning_scale, control_guidance_start,
                     control_guidance_end, callback_on_step_end_tensor_inputs):
        # the subset of pipeline_brushnet.py:573-693 that applies without tokenizer / PIL inputs
        if prompt_embeds is None:
            raise ValueError("Provide `prompt_embeds` (the text encoder stays outside this package).")
        if negative_prompt_embeds is not None and prompt_embeds.shape != negative_prompt_embeds.shape:
            raise ValueError("`prompt_embeds` and `negative_prompt_embeds` must have the same shape when passed directly, but"
                             f" got: `prompt_embeds` {prompt_embeds.shape} != `negative_prompt_embeds` {negative_prompt_embeds.shape}.")
        if not isinstance(brushnet_conditioning_scale, float):
            raise TypeError("For single brushnet: `brushnet_conditioning_scale` must be type `float`.")   # :649-650
        if control_guidance_start >= control_guidance_end:
            raise ValueError(f"control guidance start: {control_guidance_start} cannot be larger or equal to control guidance end: {control_guidance_end}.")
        if control_guidance_start < 0.0:
            raise ValueError(f"control guidance start: {control_guidance_start} can't be smaller than 0.")
        if control_guidance_end > 1.0:
            raise ValueError(f"control guidance end: {control_guidance_end} can't be larger than 1.0.")
        if callback_on_step_end_tensor_inputs is not None and any(k != "latents" for k in callback_on_step_end_tensor_inputs):
            raise ValueError("`callback_on_step_end_tensor_inputs` has to be in ['latents']")

    @torch.no_grad()
    def __call__(self, prompt=None, image=None, mask=None, depth=None, normals=None, height=None, width=None,
                 num_inference_steps: int = 50, timesteps=None, guidance_scale: float = 7.5, negative_prompt=None,
                 num_images_per_prompt: int = 1, eta: float = 0.0, generator=None, latents=None, prompt_embeds=None,
                 negative_prompt_embeds=None, ip_adapter_image=None, ip_adapter_image_embeds=None,
                 output_type: str = "latent", return_dict: bool = True,
                 cross_attention_kwargs=None, brushnet_conditioning_scale: float = 1.0, guess_mode: bool = False,
                 control_guidance_start: float = 0.0, control_guidance_end: float = 1.0, clip_skip=None,
                 callback_on_step_end=None, callback_on_step_end_tensor_inputs=("latents",), **kwargs):
        """Same parameters, in the same order, as StableDiffusionBrushNetPipeline.__call__ (pipeline_brushnet.py:848-880).
        One extra keyword rides in **kwargs: `conditioning_latents` ([b or 2b, 6, h, w], what :1188-1202 builds) for
        callers that keep the VAE outside."""
        conditioning_latents = kwargs.pop("conditioning_latents", None)
        if kwargs:
            raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            raise NotImplementedError("IP-adapter inputs are not on the MirrorFusion depth-concat path")
        if prompt is not None or negative_prompt is not None:
            raise NotImplementedError("tokenizer / CLIP are outside the hot path: pass prompt_embeds / negative_prompt_embeds")
        if timesteps is not None:
            # retrieve_timesteps (:112-121): neither scheduler of the path accepts a custom schedule in `set_timesteps`; same error
            raise ValueError(f"The current scheduler class {self.scheduler.__class__}'s `set_timesteps` does not support custom"
                             f" timestep schedules. Please check whether you are using the correct scheduler.")
        self.check_inputs(prompt_embeds, negative_prompt_embeds, brushnet_conditioning_scale, control_guidance_start,
                          control_guidance_end, callback_on_step_end_tensor_inputs)
        do_cfg = guidance_scale > 1.0                                                      # :835-836
        if not do_cfg:
            # The reference then runs the nets on the conditional batch alone and skips the combine (:1256,1310).  The fused step
            # is built for two CFG halves; with BOTH halves fed the conditional embeddings and a combine weight of exactly 1 the
            # update it computes is u + 1 (c - u) with u == c bit for bit, i.e. the conditional prediction — same result, at the
            # cost of the redundant half (guidance <= 1 is not a throughput configuration of MirrorFusion).
            negative_prompt_embeds, guidance_scale = prompt_embeds, 1.0
        if negative_prompt_embeds is None:
            # the reference encodes "" through CLIP for the unconditional half (encode_prompt, :417-446) — not a zero tensor;
            # the text encoder is outside this package, so its embedding of the empty prompt must be handed in once
            if self.empty_prompt_embeds is None:
                raise ValueError("`negative_prompt_embeds` is None: pass it, or construct the pipeline with `empty_prompt_embeds` "
                                 "(the text encoder's embedding of the empty prompt, [1 or b, 77, ctx]); zeros are not what the "
                                 "reference uses for the unconditional half")
            negative_prompt_embeds = self.empty_prompt_embeds.to(prompt_embeds).expand(prompt_embeds.shape[0], -1, -1)
        if num_images_per_prompt != 1:                                                     # encode_prompt's repeat (:403-405,447-451)
            rep = lambda t: t.repeat_interleave(num_images_per_prompt, 0)
            prompt_embeds, negative_prompt_embeds = rep(prompt_embeds), rep(negative_prompt_embeds)
            if conditioning_latents is not None and conditioning_latents.shape[0] * num_images_per_prompt == prompt_embeds.shape[0]:
                conditioning_latents = rep(conditioning_latents)
            image, mask, depth = (None if t is None else rep(t) for t in (image, mask, depth))   # prepare_image's repeat (:766)
        b = prompt_embeds.shape[0]
        ehs = torch.cat([negative_prompt_embeds, prompt_embeds])                           # uncond first (:1102-1103)
        if conditioning_latents is None:
            conditioning_latents = self._prepare_conditioning(image, mask, depth, b)
        if conditioning_latents.shape[0] == b and not guess_mode:
            conditioning_latents = torch.cat([conditioning_latents] * 2)                   # CFG duplicate (:771-772; not in guess mode)
        H, W = conditioning_latents.shape[-2:]
        if latents is None:
            latents = torch.randn(b, self.cfg.in_channels, H, W, generator=generator,
                                  device=generator.device if generator is not None else "cpu")   # randn_tensor semantics
        # guess_mode (:1076-1081,1262-1301): BrushNet on the conditional batch only, log-spaced tap scales, zeros for the uncond half
        eng = self.engine(b, H, W, guess_mode=bool(guess_mode))
        eng.set_conditioning(ehs, conditioning_latents)
        n = num_inference_steps
        keep = [1.0 - float(i / n < control_guidance_start or (i + 1) / n > control_guidance_end) for i in range(n)]   # :1236-1242
        scales = [brushnet_conditioning_scale * k for k in keep]                            # :1269-1275
        cb = None
        if callback_on_step_end is not None:
            def cb(i, t, x):
                out = callback_on_step_end(self, i, t, {"latents": x})
                return None if out is None else out.pop("latents", None)
        # eta / generator = prepare_extra_step_kwargs (:556-571): forwarded to the scheduler step only if it takes them (DDIM)
        x = eng.denoise(latents, self.scheduler, n, guidance_scale, scales, cb, eta=eta, generator=generator)
        result = x.clone()
        if output_type != "latent":
            if self.vae_decode is None:
                raise ValueError("output_type != 'latent' needs a vae_decode callable")
            result = self.vae_decode(result / 0.18215)                                      # :1342
        if not return_dict:
            return (result, None)
        return SimpleNamespace(images=result, nsfw_content_detected=None)

    def _prepare_conditioning(self, image, mask, depth, b):
        """pipeline_brushnet.py:1188-1202 given already preprocessed tensors: image [-1,1] masked RGB, mask 1-ch {0,1},
        depth [b,1,h,w] in [-1,1]; needs `vae_encode` returning the latent sample."""
        if self.vae_encode is None or image is None or mask is None or depth is None:
            raise ValueError("pass `conditioning_latents`, or `image`/`mask`/`depth` together with a `vae_encode` callable")
        lat = self.vae_encode(image) * 0.18215
        m = torch.nn.functional.interpolate(mask, size=lat.shape[-2:])
        d = torch.nn.functional.interpolate(depth, size=lat.shape[-2:])
        return torch.cat([lat, m.to(lat), d.to(lat)], 1)
