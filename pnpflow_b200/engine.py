"""UNetEngine: the velocity network v_theta(x, t) of PnP-Flow on the sm_100a engine.

Drop-in for the reference's ``model(x, t)`` call (pnpflow/methods/pnp_flow.py:19-21, pnpflow/models.py:442-495):
same signature (x fp32 [B,C,H,W] on the GPU, t fp32 [B]) and the same ``state_dict`` key scheme
(pnpflow/utils.py:225).  All arithmetic runs in libpnpflow_sm100a.so; PyTorch only owns the buffers and the stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib


def _cfg_from(obj) -> Dict:
    keys = ("input_channels", "input_height", "ch", "ch_mult", "num_res_blocks", "attn_resolutions")
    if isinstance(obj, dict):
        return {k: obj[k] for k in keys}
    return {k: getattr(obj, k) for k in keys}      # reference nn.Module (models.py:316-327) or oracle.UNetConfig


class UNetEngine:
    """``UNetEngine(model_or_cfg, state_dict=None, device='cuda', max_batch=...)``; call like the reference model."""

    def __init__(self, model_or_cfg, state_dict: Optional[Dict[str, torch.Tensor]] = None, device="cuda",
                 max_batch: int = 1, use_cuda_graph: bool = True):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pnpflow_b200 runs on CUDA devices only (no CPU fallback)")
        cfg = _cfg_from(model_or_cfg)
        if state_dict is None:
            state_dict = model_or_cfg.state_dict()
        self.cfg = cfg
        c = _lib.UNetConfigC()
        c.input_channels, c.input_height, c.ch = cfg["input_channels"], cfg["input_height"], cfg["ch"]
        c.num_levels = len(cfg["ch_mult"])
        for i, m in enumerate(cfg["ch_mult"]):
            c.ch_mult[i] = m
        c.num_res_blocks = cfg["num_res_blocks"]
        c.num_attn_resolutions = len(cfg["attn_resolutions"])
        for i, m in enumerate(cfg["attn_resolutions"]):
            c.attn_resolutions[i] = m
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pnpf_create(C.byref(c), C.byref(h)))
            self._h = h
            self.load_state_dict(state_dict)
        self.use_cuda_graph = use_cuda_graph
        self.max_batch = 0
        self._ws = None
        self._graphs = {}
        self.ensure_batch(max_batch)

    # ------------------------------------------------------------------ weights
    def expected_keys(self):
        n = self.lib.pnpf_num_weights(self._h)
        return [self.lib.pnpf_weight_name(self._h, i).decode() for i in range(n)]

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        expected = set(self.expected_keys())
        unexpected = [k for k in sd if k not in expected]
        if unexpected:
            raise RuntimeError(f"Unexpected key(s) in state_dict: {unexpected[:5]}")
        for k in expected:
            if k not in sd:
                raise RuntimeError(f"Missing key(s) in state_dict: {k}")
            w = sd[k].detach().to("cpu", torch.float32).contiguous()
            shape = (C.c_int64 * w.dim())(*w.shape)
            _lib.check(self.lib.pnpf_load_weight(self._h, k.encode(), w.data_ptr(), shape, w.dim()))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pnpf_finalize_weights(self._h))
        # the old weight arena is gone: graphs captured before this point must not be replayed any more
        self._weights_gen = getattr(self, "_weights_gen", 0) + 1
        self._graphs = {}
        if getattr(self, "_ws", None) is not None:
            mb, self.max_batch = self.max_batch, 0
            self.ensure_batch(mb)

    # ------------------------------------------------------------------ workspace
    def ensure_batch(self, batch: int):
        """Grow the workspace to hold ``batch`` images.  Graphs captured on the previous workspace stay valid: each one keeps
        its workspace tensor alive (``graphed`` stores it next to the graph) and has the old plan's pointers baked in, so a
        live PnPFlowSession is not disturbed; only NEW ``graphed()`` calls capture against the new workspace."""
        if batch <= self.max_batch:
            return
        with torch.cuda.device(self.device):
            need = self.lib.pnpf_workspace_bytes(self._h, batch)
            if need == 0:
                _lib.check(1)
            self._graphs = {}
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            base = (self._ws.data_ptr() + 1023) // 1024 * 1024
            _lib.check(self.lib.pnpf_bind_workspace(self._h, base, need, batch))
        self.max_batch = batch
        self.workspace_bytes = need

    # ------------------------------------------------------------------ forward
    def _launch(self, x, t, v, batch):
        _lib.check(self.lib.pnpf_unet_forward(self._h, x.data_ptr(), t.data_ptr(), v.data_ptr(), batch, _lib.stream_ptr()))

    def forward(self, x: torch.Tensor, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4, "x must be a CUDA fp32 [B,C,H,W] tensor"
        B = x.shape[0]
        assert list(x.shape[1:]) == [self.cfg["input_channels"], self.cfg["input_height"], self.cfg["input_height"]]
        assert t.shape == (B,)
        self.ensure_batch(B)
        x = x.contiguous()
        t = t.to(device=x.device, dtype=torch.float32).contiguous()
        v = out if out is not None else torch.empty_like(x)
        with torch.cuda.device(self.device):
            self._launch(x, t, v, B)
        return v

    __call__ = forward

    def graphed(self, batch: int):
        """Static-buffer CUDA-graph replay of one forward at a fixed batch: returns (x_buf, t_buf, v_buf, replay)."""
        if batch in self._graphs:
            return self._graphs[batch]
        self.ensure_batch(batch)
        Cc, Hh = self.cfg["input_channels"], self.cfg["input_height"]
        with torch.cuda.device(self.device):
            xb = torch.zeros(batch, Cc, Hh, Hh, device=self.device)
            tb = torch.zeros(batch, device=self.device)
            vb = torch.empty_like(xb)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):                       # warm-up: sets kernel attributes outside capture
                    self._launch(xb, tb, vb, batch)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch(xb, tb, vb, batch)
        ws, gen = self._ws, self._weights_gen     # the graph's kernels point into THIS workspace and weight arena

        def replay(_g=g, _ws=ws, _gen=gen):
            if _gen != self._weights_gen:
                raise RuntimeError("stale CUDA graph: the engine's weights were reloaded after this graph was captured; "
                                   "call UNetEngine.graphed() (or create a new PnPFlowSession) again")
            _g.replay()
        self._graphs[batch] = (xb, tb, vb, replay)
        return self._graphs[batch]

    # ------------------------------------------------------------------ introspection / debug
    @property
    def flops_per_image(self) -> float:
        return float(self.lib.pnpf_unet_flops_per_image(self._h))

    @property
    def num_launches(self) -> int:
        return int(self.lib.pnpf_unet_num_launches(self._h))

    def profile(self, batch: int):
        """One forward with CUDA events around every op.  Returns a list of dicts
        {name, kind ('tc'|'simt'), ms, flops, bytes} with flops/bytes already multiplied by ``batch``."""
        self.ensure_batch(batch)
        Cc, Hh = self.cfg["input_channels"], self.cfg["input_height"]
        names = self.op_names()
        n = len(names)
        with torch.cuda.device(self.device):
            x = torch.randn(batch, Cc, Hh, Hh, device=self.device)
            t = torch.full((batch,), 0.5, device=self.device)
            v = torch.empty_like(x)
            ms = (C.c_float * n)()
            for _ in range(2):
                _lib.check(self.lib.pnpf_profile_forward(self._h, x.data_ptr(), t.data_ptr(), v.data_ptr(), batch, ms, n,
                                                         _lib.stream_ptr()))
        out = []
        for i, name in enumerate(names):
            kind, fl, by = C.c_int(), C.c_double(), C.c_double()
            _lib.check(self.lib.pnpf_debug_op_info(self._h, i, C.byref(kind), C.byref(fl), C.byref(by)))
            out.append(dict(name=name, kind="tc" if kind.value == 1 else "simt", ms=float(ms[i]),
                            flops=fl.value * batch, bytes=by.value * batch, impl=self.lib.pnpf_debug_op_impl(self._h, i).decode()))
        return out

    def op_names(self):
        n = self.lib.pnpf_debug_num_ops(self._h)
        return [self.lib.pnpf_debug_op_name(self._h, i).decode() for i in range(n)]

    def debug_activation(self, x, t, op_index: int) -> torch.Tensor:
        """Run ops [0, op_index] and return op_index's output as fp32 NCHW (parity tests)."""
        B = x.shape[0]
        self.ensure_batch(B)
        x = x.contiguous()
        t = t.to(device=x.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pnpf_debug_forward_partial(self._h, x.data_ptr(), t.data_ptr(), B, op_index + 1, _lib.stream_ptr()))
            dims = (C.c_int * 3)()
            # first call only reports the dims (it fails with "destination too small" after filling them)
            probe = torch.empty(1, device=self.device)
            rc = self.lib.pnpf_debug_read_op_output(self._h, op_index, B, probe.data_ptr(), 0, C.byref(dims), _lib.stream_ptr())
            n = B * dims[0] * dims[1] * dims[2]
            if n == 0:
                _lib.check(rc)
            buf = torch.empty(n, device=self.device)
            _lib.check(self.lib.pnpf_debug_read_op_output(self._h, op_index, B, buf.data_ptr(), n, C.byref(dims), _lib.stream_ptr()))
        return buf.view(B, dims[0], dims[1], dims[2])

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.pnpf_destroy(self._h)
                self._h = None
        except Exception:
            pass
