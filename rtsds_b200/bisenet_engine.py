"""Execution plan for BiSeNet (reference models/bisenet/build_bisenet.py:141-172).

A plan is built once per (batch, height, width, mode, precision): it owns the
NHWC activation buffers, the packed bf16 weights, the folded BatchNorm vectors
and an ordered list of kernel launches into librtsds_b200.so.  PyTorch supplies
memory, streams and CUDA-graph capture; no torch operator computes anything on
the path.

Data layout in HBM (bf16 production mode; fp32 in check mode):
  image        NCHW fp32   [N,3,H,W]          (API boundary, read by both stems)
  sp1/cp0      NHWC        [N,H/2,W/2,64]
  sp2          NHWC        [N,H/4,W/4,128]
  cat          NHWC        [N,H/8,W/8,1024]   sx | gated+resized cx1 | cx2 slots
                                              (torch.cat of :153/:72 never happens)
  layer1..4    NHWC        ResNet-18 stage buffers (tmp / ping / pong / downsample)
  feat         NHWC fp32   [N,H/8,W/8,32]     FFM ConvBlock output (19 valid channels)
  z, z1, z2    NHWC fp32   [N,H/8,W/8,32]     logits at 1/8 resolution
  logits       NCHW fp32   [N,19,H,W]         (API boundary)
"""
from __future__ import annotations

import os

import torch

from . import ops, weights_epoch
from .ops import ACT_NONE, ACT_RELU, BF16, F16, F32


class _BN:
    """Folded (eval) or batch-statistics (train) BatchNorm attached to a conv."""

    def __init__(self, plan, bn, c):
        self.bn = bn
        self.c = c
        dev = plan.device
        self.scale = torch.empty(c, dtype=torch.float32, device=dev)
        self.shift = torch.empty(c, dtype=torch.float32, device=dev)
        if plan.train:
            self.stats = plan.alloc_stats(c)
            self.save_mean = torch.empty(c, dtype=torch.float32, device=dev)
            self.save_invstd = torch.empty(c, dtype=torch.float32, device=dev)


class BiSeNetPlan:
    def __init__(self, model, n: int, h: int, w: int, train: bool, precision: str = "bf16"):
        p0 = model.conv.weight
        if not p0.is_cuda and not ops._lib.dry_run():
            raise ops._lib.RtsdsError("BiSeNet parameters must live on a CUDA device (no CPU fallback)")
        ops.check(ops.lib().rtsds_check_device(), "device check")
        self.model = model
        self.device = p0.device
        self.n, self.h, self.w = n, h, w
        self.train = train
        self.precision = precision
        if precision not in ("bf16", "fp16", "fp32", "bf16_simt"):
            raise ops._lib.RtsdsError(f"unknown precision {precision!r}")
        if precision == "fp16" and train:
            raise ops._lib.RtsdsError("fp16 is the eval-mode (inference) precision; training runs in bf16 or fp32")
        self.dt = F32 if precision == "fp32" else F16 if precision == "fp16" else BF16
        self.use_tc = precision in ("bf16", "fp16")  # "bf16_simt": bf16 storage, CUDA-core convs (cross-check)
        self.tdt = ops.torch_dtype(self.dt)
        self.nc = model.conv.weight.shape[0]
        if model._context_name not in ("resnet18", "resnet101"):
            raise ops._lib.RtsdsError(f"context paths: resnet18 and resnet101 (build_bisenet.py:95-113); got {model._context_name!r}")
        if self.nc > 32:
            raise ops._lib.RtsdsError("num_classes > 32 is not supported by the fused head kernels")
        self._stats_chunks = []
        self._keep = []
        self._stats_total = 0
        self.pre_steps = []      # read the caller's input tensor (outside the CUDA graph)
        self.steps = []          # everything between the stems and the low-res logits
        self.head_steps = []     # first nodes of the graph, before the spatial-path branch forks (the s2d stem GEMM)
        self.pack_steps = []     # weight repack / BN fold (run when parameters change)
        self.ws = None
        self.x_f32 = None
        self.ws_ds = None
        self.side_ds = None
        self._ws_bytes = 0
        self._param_version = None
        self._n_state = -1
        self.generation = 0
        self.graph = None
        self.side = None
        self.ws_side = None
        self._side_branch = False
        self.n_sp_steps = self.n_join = 0
        self._build()

    # ------------------------------------------------------------------ helpers
    def buf(self, *shape, dtype=None):
        # kernels are handed raw pointers, so the plan must keep every buffer alive itself
        t = torch.empty(shape, dtype=self.tdt if dtype is None else dtype, device=self.device)
        self._keep.append(t)
        return t

    def zeros(self, *shape, dtype=None):
        t = torch.zeros(shape, dtype=self.tdt if dtype is None else dtype, device=self.device)
        self._keep.append(t)
        return t

    def alloc_stats(self, c):
        off = self._stats_total
        self._stats_total += 2 * c
        self._stats_chunks.append((off, c))
        return off

    def _stats_view(self, off, c):
        return self.stats_all[off:off + 2 * c]

    def _need_ws(self, d):
        if self.use_tc:
            need = int(ops.lib().rtsds_conv2d_tc_workspace_bytes(d))
            self._ws_bytes = max(self._ws_bytes, need)

    def note_ws(self, b):
        self._ws_bytes = max(self._ws_bytes, b)

    def _conv_tapn(self, conv, bnmod, x, xshape, y, y_ld, act, in_ld, in_scale=None, gap_out=None):
        """Skinny-output k x k conv + folded BatchNorm + activation in the taps-as-N form (rtsds_b200/tapn.py); y fp32.
        in_scale = (c0, c1, factor): input channels [c0, c1) of x are stored divided by `factor` (block exponent)."""
        from .tapn import TapNConv

        cout = conv.weight.shape[0]
        tn = TapNConv(self, conv, x, xshape, in_ld, train=False, in_scale=in_scale)
        bn = _BN(self, bnmod, cout)
        self.pack_steps.append(lambda: ops.bn_fold(bnmod, bn.scale, bn.shift, conv.bias))
        self.steps.append(lambda: tn.forward(bn.scale, bn.shift, act, None, y, y_ld, gap_out))

    def _conv(self, conv, bnmod, x, xshape, y, out_ld, act, *, in_ld=None, residual=None, res_ld=0, out_dtype=None,
              x_off=0, y_off=0, bias=None, steps=None, in_scale=None, gap_out=None):
        """Append conv (+BN/bias, +residual, +act) reading NHWC x -> NHWC y.  Returns (oh, ow)."""
        steps = self.steps if steps is None else steps
        n, h, w, cin = xshape
        in_ld = cin if in_ld is None else in_ld
        cout = conv.weight.shape[0]
        k = conv.kernel_size[0]
        out_dtype = self.dt if out_dtype is None else out_dtype
        d = ops.make_conv_desc(n, h, w, cin, in_ld, cout, out_ld, k, conv.stride[0], conv.padding[0], conv.dilation[0],
                               act=ACT_NONE, in_dtype=self.dt, out_dtype=out_dtype, res_ld=res_ld)
        wpk = self.buf(ops.cout_pad(cout), k * k, cin)
        self.pack_steps.append(lambda: ops.pack_conv_weight(conv.weight, self.dt, wpk))
        if in_scale is not None:
            c0_, c1_, fac_ = in_scale
            self.pack_steps.append(lambda: ops.scale_packed_channels(wpk, wpk.shape[0] * wpk.shape[1], cin, c0_, c1_, fac_))
        xp = x.data_ptr() + x_off * x.element_size()
        yp = y.data_ptr() + y_off * y.element_size()
        rp = residual.data_ptr() if residual is not None else None
        self._need_ws(d)
        use_tc = self.use_tc
        n_pix = n * d.oh * d.ow
        side = self._side_branch          # convs of the side-stream branch get their own split-K workspace

        if gap_out is not None:      # eval: AdaptiveAvgPool2d(1) of this layer's output leaves its epilogue as per-CTA partial means
            gap_out["parts"] = ops.conv2d_tc_gap_parts(d)
            gap_out["buf"] = self.buf(n, gap_out["parts"], cout, dtype=torch.float32)

        def launch(desc, scale, shift, res, stats):
            if use_tc and gap_out is not None:
                ops.conv2d_tc_gap(desc, xp, wpk, yp, scale, shift, res, gap_out["buf"], self.ws)
            elif use_tc:
                ops.conv2d_tc(desc, xp, wpk, yp, scale, shift, res, stats,
                              self.ws_ds if side == 2 else (self.ws_side if side else self.ws))
            else:
                ops.conv2d_simt(desc, xp, wpk, yp, scale, shift, res, stats)

        if bnmod is None:
            d.act = act
            b = conv.bias if bias is None else bias
            steps.append(lambda: launch(d, None, b.detach() if b is not None else None, rp, None))
        elif not self.train:
            bn = _BN(self, bnmod, cout)
            d.act = act
            self.pack_steps.append(lambda: ops.bn_fold(bnmod, bn.scale, bn.shift, conv.bias))
            steps.append(lambda: launch(d, bn.scale, bn.shift, rp, None))
        else:
            bn = _BN(self, bnmod, cout)

            def run_train():
                st = self._stats_view(bn.stats, cout)
                launch(d, None, None, None, st)              # raw conv output + per-channel sums
                ops.bn_finalize_apply_ptr(st, n_pix, bnmod, bn.scale, bn.shift, bn.save_mean, bn.save_invstd, yp, yp, n_pix, cout,
                                          rp, act, 0.0, out_ld, out_ld, res_ld if res_ld else cout, out_dtype, out_dtype)

            steps.append(run_train)
        return d.oh, d.ow

    # ------------------------------------------------------------------ plan construction
    def _build(self):
        m = self.model
        n, H, W = self.n, self.h, self.w
        cs = ops.conv_out_size
        nc = self.nc
        f32 = torch.float32

        # ---- spatial path (reference :21-32) ----
        sp = m.saptial_path
        h2, w2 = cs(H, 3, 2, 1), cs(W, 3, 2, 1)
        h4, w4 = cs(h2, 3, 2, 1), cs(w2, 3, 2, 1)
        h8, w8 = cs(h4, 3, 2, 1), cs(w4, 3, 2, 1)
        self.h8, self.w8 = h8, w8
        import os

        fused_stems = self.use_tc and not self.train
        # eval stems: "gather" = the thread-gathered im2col kernel of csrc/stem_tc.cu (default: 34 us at b=1 512x1024);
        # "s2d" = space-to-depth pack + ONE TMA-fed tcgen05 GEMM for both stems (N = 64 + 64, K = 4 taps x 64 = 256 instead of
        # 147): measured 9 us SLOWER per frame (pack + a GEMM with 74 % more K), kept as an A/B switch (RTSDS_STEM=s2d)
        self.stem_mode = os.environ.get("RTSDS_STEM", "gather") if fused_stems else "direct"
        s2d = self.stem_mode == "s2d"
        stems = self.buf(n, h2, w2, 128) if s2d else None       # context-path map in channels 0..63, spatial-path map in 64..127
        sp1 = stems if s2d else self.buf(n, h2, w2, 64)
        sp2 = self.buf(n, h4, w4, 128)
        # concat buffer of build_bisenet.py:153,72: 256 spatial-path channels | cx1 | cx2 (1024 wide for resnet18, 3328 for resnet101)
        ccat = m.feature_fusion_module.convblock.conv1.weight.shape[1]
        cat = self.buf(n, h8, w8, ccat)
        self.cat, self.ccat = cat, ccat
        if not fused_stems:
            self._stem(sp.convblock1.conv1, sp.convblock1.bn, sp1, 3, 2, 1)
        self._side_branch = True
        if s2d:
            self._conv(sp.convblock2.conv1, sp.convblock2.bn, sp1, (n, h2, w2, 64), sp2, 128, ACT_RELU, in_ld=128, x_off=64)
        else:
            self._conv(sp.convblock2.conv1, sp.convblock2.bn, sp1, (n, h2, w2, 64), sp2, 128, ACT_RELU)
        self._conv(sp.convblock3.conv1, sp.convblock3.bn, sp2, (n, h4, w4, 128), cat, ccat, ACT_RELU)
        self._side_branch = False
        # the spatial path is independent of the context path until the concat buffer is consumed: its two convs run on a
        # side stream (a parallel branch of the CUDA graph) and fill SMs the small layer-3/4 grids leave idle
        self.n_sp_steps = len(self.steps)

        # ---- context path: ResNet-18 (build_contextpath.py:18-29) ----
        cp = m.context_path
        ch2, cw2 = cs(H, 7, 2, 3), cs(W, 7, 2, 3)
        ph, pw = ops.maxpool_out_size(ch2), ops.maxpool_out_size(cw2)
        # RTSDS_STEM_POOL=1 (eval): the max-pool rides on the stem kernel's epilogue (16-byte max-reductions into a zeroed pooled
        # map; the 1/2-resolution context-path map is never written): 27 -> 26 launches and 34 MB less traffic per frame, but
        # measured time-neutral at b=1 (0.3433 vs 0.3417 ms: the reductions cost the stem what the pool launch saved) — off
        # by default, the separate max-pool launch stays.
        self.pool_in_stem = fused_stems and not s2d and not self.train and os.environ.get("RTSDS_STEM_POOL", "0") == "1"
        cp0 = stems if s2d else (None if self.pool_in_stem else self.buf(n, ch2, cw2, 64))
        x = self.zeros(n, ph, pw, 64) if self.pool_in_stem else self.buf(n, ph, pw, 64)
        if s2d:
            self._stem_pair_s2d(cp.conv1, cp.bn1, sp.convblock1.conv1, sp.convblock1.bn, stems)
        elif fused_stems:
            self._stem_pair(cp.conv1, cp.bn1, sp.convblock1.conv1, sp.convblock1.bn, cp0, sp1, pool=x if self.pool_in_stem else None)
        else:
            self._stem(cp.conv1, cp.bn1, cp0, 7, 2, 3)
        if self.pool_in_stem:
            pass
        elif s2d:
            self.steps.append(lambda cp0=cp0, x=x: ops.maxpool3x3s2_ld(cp0, n, ch2, cw2, 64, 128, self.dt, x))
        else:
            self.steps.append(lambda cp0=cp0, x=x: ops.maxpool3x3s2(cp0, x))
        pool_buf = x
        shape = (n, ph, pw, 64)
        feats = []
        # eval-mode tensor-core path: the global average pools of feature3 / feature4 (ARM pooling, context-path tail) are
        # accumulated by the epilogue of the conv that produces them; ARM gates + gated resizes are one launch
        self.fused_arm = self.use_tc and not self.train
        self._gap_for = {}
        if self.fused_arm:
            self._gap3, self._gap4 = {}, {}          # filled by _conv: per-CTA partial means [n][parts][c] (deterministic)
            self._gap_for[id(cp.layer3[-1])] = self._gap3
            self._gap_for[id(cp.layer4[-1])] = self._gap4
        for layer in (cp.layer1, cp.layer2, cp.layer3, cp.layer4):
            for bi, blk in enumerate(layer):
                x, shape = (self._bottleneck if hasattr(blk, "conv3") else self._basic_block)(blk, x, shape)
                if self.pool_in_stem and layer is cp.layer1 and bi == 0:
                    # the pooled map has been consumed (conv input + shortcut of the first block): clear it for the next frame's
                    # max-reductions on the side branch; the next ("join",) — the first downsample block — orders it
                    self.steps.append(("fork", [lambda pool_buf=pool_buf: pool_buf.zero_()]))
            feats.append((x, shape))
        if self.pool_in_stem:
            self.steps.append(("join",))
        (f3, s3), (f4, s4) = feats[2], feats[3]
        if min(h8, w8, s4[1], s4[2]) <= 0:
            raise ops._lib.RtsdsError("input too small for BiSeNet")

        # ---- ARMs, tail, gated resize into the concat buffer (reference :147-153) ----
        arm1, arm2 = m.attention_refinement_module1, m.attention_refinement_module2
        c3, c4 = s3[3], s4[3]
        if 256 + c3 + c4 != ccat:
            raise ops._lib.RtsdsError(f"feature fusion module expects {ccat} input channels, context path gives 256+{c3}+{c4}")
        pooled3 = self._gap3["buf"] if self.fused_arm else self.buf(n, c3, dtype=f32)
        pooled4 = self._gap4["buf"] if self.fused_arm else self.buf(n, c4, dtype=f32)
        gate3 = self.buf(n, c3, dtype=f32)
        gate4 = self.buf(n, c4, dtype=f32)
        self.arm_saved = dict(pooled3=pooled3, pooled4=pooled4, gate3=gate3, gate4=gate4)
        if self.train:
            for k_, c_ in (("lin3", c3), ("xhat3", c3), ("lin4", c4), ("xhat4", c4)):
                self.arm_saved[k_] = self.buf(n, c_, dtype=f32)
        sv = self.arm_saved
        tr = self.train
        dt = self.dt
        fused_arm = self.fused_arm
        if fused_arm and (c3 % 32 or c4 % 32):
            raise ops._lib.RtsdsError("context-path channel counts must be multiples of 32")
        if not fused_arm:
            self.steps.append(lambda: ops.global_avgpool(f3, n, s3[1] * s3[2], c3, c3, pooled3))
            self.steps.append(lambda: ops.global_avgpool(f4, n, s4[1] * s4[2], c4, c4, pooled4))
            self.steps.append(lambda: ops.arm_gate(pooled3, arm1.conv, arm1.bn, tr, n, c3, gate3, None, sv.get("lin3"), sv.get("xhat3")))
            # cx2 = ARM2(cx2) * tail, tail = GAP(feature4) = pooled4 (build_contextpath.py:27-28)
            self.steps.append(lambda: ops.arm_gate(pooled4, arm2.conv, arm2.bn, tr, n, c4, gate4, pooled4, sv.get("lin4"), sv.get("xhat4")))
            self.steps.append(lambda: ops.gate_resize_nhwc(f3, n, s3[1], s3[2], c3, c3, gate3, h8, w8, cat, ccat, 256, dt))
        # fp16 inference: `cx2 * tail` (reference :149) is QUADRATIC in the activations (feature4 times its own global
        # mean), the one tensor of the net whose range can leave fp16's 65504 while everything linear stays tame; its slot
        # of the concat buffer carries a block exponent of 2^-8, undone in the FFM conv's weights for those channels
        self.cx2_scale = 2.0 ** -8 if self.dt == F16 else 1.0
        cx2s = self.cx2_scale
        if fused_arm:
            self.steps.append(lambda: ops.arm_gate_resize(
                ops.arm_side(f3, pooled3, arm1, s3[1], s3[2], c3, 256, parts=self._gap3["parts"]),
                ops.arm_side(f4, pooled4, arm2, s4[1], s4[2], c4, 256 + c3, mul_pooled=True, out_scale=cx2s, parts=self._gap4["parts"]),
                dt, n, h8, w8, cat, ccat))
        else:
            self.steps.append(lambda: ops.gate_resize_nhwc(f4, n, s4[1], s4[2], c4, c4, gate4, h8, w8, cat, ccat, 256 + c3, dt, cx2s))
        self.f3, self.s3, self.f4, self.s4 = f3, s3, f4, s4
        self.n_join = len(self.steps)        # everything from here on reads the spatial-path slot of the concat buffer

        # ---- auxiliary heads, train only (reference :155-159) ----
        if self.train:
            self.z1 = self.buf(n, h8, w8, 32, dtype=f32)
            self.z2 = self.buf(n, h8, w8, 32, dtype=f32)
            self._conv(m.supervision1, None, cat, (n, h8, w8, c3), self.z1, 32, ACT_NONE, in_ld=ccat, x_off=256, out_dtype=F32)
            self._conv(m.supervision2, None, cat, (n, h8, w8, c4), self.z2, 32, ACT_NONE, in_ld=ccat, x_off=256 + c3, out_dtype=F32)

        # ---- feature fusion module + final 1x1 conv at 1/8 resolution (reference :162-167) ----
        ffm = m.feature_fusion_module
        self.feat = self.buf(n, h8, w8, 32, dtype=f32)
        self.pooled_f = self.buf(n, nc, dtype=f32)
        self.pooled_f_parts = 1
        self.attn = self.buf(n, nc, dtype=f32)
        self.z = self.buf(n, h8, w8, 32, dtype=f32)
        from . import tapn

        in_scale = (256 + c3, ccat, 1.0 / self.cx2_scale) if self.cx2_scale != 1.0 else None
        feat, pooled_f, z, attn = self.feat, self.pooled_f, self.z, self.attn
        final = m.conv if m.with_interpolation else None
        # eval with the x8 head: the FFM attention, the final 1x1 conv and the resize to the NCHW logits are ONE kernel that
        # runs when the caller asks for the logits (logits()); z at 1/8 resolution is never written
        self.fused_tail = (not self.train) and final is not None
        if self.fused_tail and tapn.applicable(ffm.convblock.conv1):
            # the FFM feature's AdaptiveAvgPool2d(1) (:75) leaves the gather as one partial mean per block
            self.pooled_f_parts = int(ops.lib().rtsds_tapn_gather_parts(n, h8, w8))
            pooled_f = self.pooled_f = self.buf(n, self.pooled_f_parts, nc, dtype=f32)
            self._conv_tapn(ffm.convblock.conv1, ffm.convblock.bn, cat, (n, h8, w8, ccat), self.feat, 32, ACT_RELU, ccat,
                            in_scale=in_scale, gap_out=pooled_f)
        elif not self.train and tapn.applicable(ffm.convblock.conv1):
            self._conv_tapn(ffm.convblock.conv1, ffm.convblock.bn, cat, (n, h8, w8, ccat), self.feat, 32, ACT_RELU, ccat,
                            in_scale=in_scale)
            self.steps.append(lambda: ops.global_avgpool(feat, n, h8 * w8, nc, 32, pooled_f))
        else:
            self._conv(ffm.convblock.conv1, ffm.convblock.bn, cat, (n, h8, w8, ccat), self.feat, 32, ACT_RELU, out_dtype=F32,
                       in_scale=in_scale)
            self.steps.append(lambda: ops.global_avgpool(feat, n, h8 * w8, nc, 32, pooled_f))
        if not self.fused_tail:
            self.steps.append(lambda: ops.ffm_head(feat, F32, 32, pooled_f, n, h8 * w8, nc, ffm.conv1, ffm.conv2, final, z, 32, attn))

        self.stats_all = torch.zeros(max(self._stats_total, 1), dtype=f32, device=self.device)
        if self._ws_bytes:
            self.ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
            self.ws_side = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)
            self.ws_ds = torch.empty(self._ws_bytes, dtype=torch.uint8, device=self.device)

    def _stem_pair_s2d(self, conv7, bn7, conv3, bn3, y128):
        """Both stems as ONE 4-tap implicit GEMM over the padded space-to-depth image P (csrc/conv_tc.cu rtsds_stem_s2d_*):
        the 3x3 s2 p1 window is the centre of the 7x7 s2 p3 window, so both filters live in the same virtual
        [cout, 64, 4, 1] weight layout and share every A tile.  The pack kernel reads the caller's image — fp32, or the RAW
        uint8 frame with transforms.Normalize folded in — and runs eagerly; the GEMM reads only plan-owned
        memory and is the first node of the CUDA graph."""
        from .input_pipeline import stem_affine

        n = self.n
        oh, ow, pshape = ops.stem_s2d_shape(n, self.h, self.w)
        P = self.buf(*pshape)
        w2 = self.buf(128, 64, 4, 1, dtype=torch.float32)
        wpk = self.buf(128, 4, 64)
        scale = self.buf(128, dtype=torch.float32)
        shift = self.buf(128, dtype=torch.float32)
        self.pack_steps.append(lambda: (ops.stem_s2d_weight(conv7.weight, w2[:64]), ops.stem_s2d_weight(conv3.weight, w2[64:]),
                                        ops.pack_conv_weight(w2, self.dt, wpk)))
        self.pack_steps.append(lambda: ops.bn_fold(bn7, scale[:64], shift[:64]))
        self.pack_steps.append(lambda: ops.bn_fold(bn3, scale[64:], shift[64:]))

        def pack(x):
            if x.dtype == torch.uint8:
                sc, bi = stem_affine(self.model)
                ops.stem_s2d_pack_ex(x, P, sc, bi)
            else:
                ops.stem_s2d_pack_ex(x, P)

        self.pre_steps.append(pack)
        self.head_steps.append(lambda: ops.stem_s2d_conv_fwd_dt(P, n, oh, ow, wpk, 128, y128, 128, self.dt, scale, shift, ACT_RELU))

    def _stem_pair(self, conv7, bn7, conv3, bn3, y_cp, y_sp, pool=None):
        """Both stems in one tensor-core kernel (csrc/stem_tc.cu); eval mode: BN folded + ReLU."""
        wpk = self.buf(128, 192)
        scale = self.buf(128, dtype=torch.float32)
        shift = self.buf(128, dtype=torch.float32)
        self.pack_steps.append(lambda: ops.stem_pack_weights(conv7.weight, conv3.weight, wpk))
        self.pack_steps.append(lambda: ops.bn_fold(bn7, scale[:64], shift[:64]))
        self.pack_steps.append(lambda: ops.bn_fold(bn3, scale[64:], shift[64:]))
        from .input_pipeline import stem_affine

        def run(x):
            if x.dtype == torch.uint8:
                # raw uint8 frame: convert + normalise into a plan-owned fp32 image (one 3 us pass, csrc/input.cu)
                if self.x_f32 is None:
                    self.x_f32 = self.buf(self.n, 3, self.h, self.w, dtype=torch.float32)
                sc, bi = stem_affine(self.model)
                ops.check(ops.lib().rtsds_image_u8_to_f32(ops._p(x), self.n, 3, self.h, self.w, self.h, self.w, sc, bi, ops._p(self.x_f32),
                                                          ops._s()), "image_u8_to_f32")
                x = self.x_f32
            if pool is not None:
                ops.stem_pair_tc_fwd_pool(x, wpk, pool, y_sp, scale, shift)
            else:
                ops.stem_pair_tc_fwd(x, wpk, y_cp, y_sp, scale, shift, True)

        self.pre_steps.append(run)

    def _stem(self, conv, bnmod, y, k, stride, pad):
        cout = conv.weight.shape[0]
        if not self.train:
            bn = _BN(self, bnmod, cout)
            self.pack_steps.append(lambda: ops.bn_fold(bnmod, bn.scale, bn.shift))
            self.pre_steps.append(lambda x: ops.stem_conv(x, conv.weight, y, k, stride, pad, bn.scale, bn.shift, ACT_RELU))
        else:
            bn = _BN(self, bnmod, cout)
            n_pix = y.shape[0] * y.shape[1] * y.shape[2]

            def run(x):
                st = self._stats_view(bn.stats, cout)
                ops.stem_conv(x, conv.weight, y, k, stride, pad, stats=st)
                ops.bn_finalize(st, n_pix, bnmod, bn.scale, bn.shift, bn.save_mean, bn.save_invstd)
                ops.scale_shift_act(y, y, n_pix, cout, bn.scale, bn.shift, None, ACT_RELU)

            self.pre_steps.append(run)

    def _basic_block(self, blk, x, shape):
        """torchvision BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut(x))."""
        n, h, w, cin = shape
        cout = blk.conv1.weight.shape[0]
        st = blk.conv1.stride[0]
        oh, ow = ops.conv_out_size(h, 3, st, 1), ops.conv_out_size(w, 3, st, 1)
        t = self.buf(n, oh, ow, cout)
        y = self.buf(n, oh, ow, cout)
        if blk.downsample is not None:
            # the 1x1 stride-2 shortcut conv only depends on the block input: it runs on its own stream (a parallel branch of
            # the CUDA graph) beside conv1 instead of between conv1 and conv2 — 4.5-5.4 us each off the critical path at b=1
            ds = self.buf(n, oh, ow, cout)
            branch = []
            self._side_branch = 2
            self._conv(blk.downsample[0], blk.downsample[1], x, shape, ds, cout, ACT_NONE, steps=branch)
            self._side_branch = False
            self.steps.append(("fork", branch))
            self._conv(blk.conv1, blk.bn1, x, shape, t, cout, ACT_RELU)
            self.steps.append(("join",))
            res = ds
        else:
            self._conv(blk.conv1, blk.bn1, x, shape, t, cout, ACT_RELU)
            res = x
        self._conv(blk.conv2, blk.bn2, t, (n, oh, ow, cout), y, cout, ACT_RELU, residual=res, res_ld=cout,
                   gap_out=self._gap_for.get(id(blk)))
        return y, (n, oh, ow, cout)

    def _bottleneck(self, blk, x, shape):
        """torchvision Bottleneck (resnet101 context path, build_contextpath.py:32-56): 1x1 -> 3x3 (carries the stride) -> 1x1,
        each with folded BatchNorm; relu(out + shortcut(x))."""
        n, h, w, cin = shape
        planes = blk.conv1.weight.shape[0]
        cout = blk.conv3.weight.shape[0]
        st = blk.conv2.stride[0]
        oh, ow = ops.conv_out_size(h, 3, st, 1), ops.conv_out_size(w, 3, st, 1)
        t1 = self.buf(n, h, w, planes)
        t2 = self.buf(n, oh, ow, planes)
        y = self.buf(n, oh, ow, cout)
        self._conv(blk.conv1, blk.bn1, x, shape, t1, planes, ACT_RELU)
        self._conv(blk.conv2, blk.bn2, t1, (n, h, w, planes), t2, planes, ACT_RELU)
        if blk.downsample is not None:
            ds = self.buf(n, oh, ow, cout)
            self._conv(blk.downsample[0], blk.downsample[1], x, shape, ds, cout, ACT_NONE)
            res = ds
        else:
            res = x
        self._conv(blk.conv3, blk.bn3, t2, (n, oh, ow, planes), y, cout, ACT_RELU, residual=res, res_ld=cout,
                   gap_out=self._gap_for.get(id(blk)))
        return y, (n, oh, ow, cout)

    # ------------------------------------------------------------------ execution
    def _params_version(self):
        ts = self.__dict__.get("_version_tensors")
        if ts is None or len(ts) != self._n_state:           # (re)collect on first use; module surgery changes the count
            ts = list(self.model.parameters()) + list(self.model.buffers())
            self._version_tensors, self._n_state = ts, len(ts)
        return (sum([t._version for t in ts]), ts[0].data_ptr(), self.model.conv.weight.data_ptr(), weights_epoch.value())

    def refresh_weights(self, force=False):
        ver = self._params_version()
        if force or ver != self._param_version:
            for s in self.pack_steps:
                s()
            self._param_version = self._params_version()
            return True
        return False

    def run_pre(self, x):
        for s in self.pre_steps:
            s(x)

    def _run(self, steps, main, parallel):
        """Plain steps are callables; ("fork", [steps]) runs its steps on the shortcut stream from this point on,
        ("join",) makes the main stream wait for them (graph capture turns both into dependency edges)."""
        for s in steps:
            if not isinstance(s, tuple):
                s()
            elif s[0] == "fork":
                if parallel:
                    self.side_ds.wait_stream(main)
                    with torch.cuda.stream(self.side_ds):
                        for b in s[1]:
                            b()
                else:
                    for b in s[1]:
                        b()
            elif parallel:
                main.wait_stream(self.side_ds)

    def run_mid(self):
        for s in self.head_steps:
            s()
        if ops._lib.dry_run() or self.n_sp_steps == 0:
            self._run(self.steps, None, False)
            return
        main = torch.cuda.current_stream(self.device)
        if self.side is None:
            self.side = torch.cuda.Stream(self.device)
            self.side_ds = torch.cuda.Stream(self.device)
        self.side.wait_stream(main)                           # fork
        with torch.cuda.stream(self.side):
            self._run(self.steps[:self.n_sp_steps], None, False)
        self._run(self.steps[self.n_sp_steps:self.n_join], main, True)
        main.wait_stream(self.side)                           # join
        self._run(self.steps[self.n_join:], main, True)

    def forward_lowres(self, x, use_graph: bool):
        """Run everything up to the 1/8-resolution logits (self.z [, z1, z2])."""
        # Scanning the version counters of ~200 tensors costs ~20 us of Python: at batch 1 that is 5 % of a frame and sits
        # BEFORE the first launch.  The O(1) part of the key (the process-wide weights epoch, bumped by every backward
        # pass) is checked up front; the per-tensor scan runs after the frame has been enqueued, overlapping the GPU,
        # and in the rare case it finds an in-place edit the operands are re-packed and the frame is issued again.
        late_check = not self.train and self._param_version is not None and self._param_version[-1] == weights_epoch.value()
        if not late_check:
            self.refresh_weights()
        self._enqueue(x, use_graph)
        if late_check and self.refresh_weights():
            self._enqueue(x, use_graph)

    def _enqueue(self, x, use_graph: bool):
        if self.train and self._stats_total:
            self.stats_all.zero_()
        self.generation += 1
        self.run_pre(x)
        if use_graph and not self.train and not ops._lib.dry_run():
            if self.graph is None:
                self.run_mid()                       # warm-up: sets kernel attributes, primes caches
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.run_mid()
                self.graph = g
                if getattr(self, "pool_in_stem", False):
                    self.run_pre(x)                  # the warm-up pass consumed (and cleared) the stem's pooled map: refill it
            self.graph.replay()
        else:
            self.run_mid()

    def logits(self, z):
        n, H, W = self.n, self.h, self.w
        if self.model.with_interpolation:
            # F.interpolate(scale_factor=8) (reference :166): output is 8x the 1/8-resolution map
            out = torch.empty((n, self.nc, self.h8 * 8, self.w8 * 8), dtype=torch.float32, device=self.device)
            if self.fused_tail and z is self.z:
                ffm = self.model.feature_fusion_module
                ops.ffm_head_resize(self.feat, 32, self.pooled_f, n, self.h8, self.w8, self.nc, ffm.conv1, ffm.conv2, self.model.conv,
                                    out, self.attn, self.pooled_f_parts)
                return out
        else:
            out = torch.empty((n, self.nc, self.h8, self.w8), dtype=torch.float32, device=self.device)
        ops.resize_to_nchw(z, n, self.h8, self.w8, self.nc, 32, out)
        return out

    def logits_aux(self, z):
        """Auxiliary heads are resized to the INPUT size (reference :158-159), not x8."""
        out = torch.empty((self.n, self.nc, self.h, self.w), dtype=torch.float32, device=self.device)
        ops.resize_to_nchw(z, self.n, self.h8, self.w8, self.nc, 32, out)
        return out


def eval_precision(model) -> str:
    """Precision of the eval-mode forward: the production mode ("bf16") runs inference with fp16 operands — the same
    tcgen05 kind::f16 rate, 3 more significand bits: on BASELINE config 1 the argmax agrees with the fp32 reference on
    99.98 % of the pixels (bf16: 99.78 %, below north_star's 99.9 %).  model.rtsds_eval_precision = "bf16" forces bf16."""
    prec = model.rtsds_precision
    if prec == "bf16":
        prec = getattr(model, "rtsds_eval_precision", "fp16")
    return prec


def _get_plan(model, x, train):
    plans = model.__dict__.setdefault("_rtsds_plans", {})
    n, _, h, w = x.shape
    # lane: independent plan instances (own buffers / CUDA graph) so that several frames can be in flight on
    # different streams (rtsds_b200/serving.py)
    prec = model.rtsds_precision if train else eval_precision(model)
    key = (n, h, w, bool(train), prec, x.device.index, int(getattr(model, "rtsds_lane", 0)))
    plan = plans.get(key)
    if plan is None:
        plan = BiSeNetPlan(model, n, h, w, bool(train), prec)
        plans[key] = plan
    return plan


def bisenet_forward(model, x):
    """BiSeNet.forward (reference :141-172): train -> (result, cx1_sup, cx2_sup); eval -> result."""
    if not x.is_cuda and not ops._lib.dry_run():
        raise ops._lib.RtsdsError("BiSeNet.forward needs a CUDA tensor: rtsds_b200 has no CPU fallback")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected input [N,3,H,W], got {tuple(x.shape)}")
    if x.dtype == torch.uint8:
        # raw uint8 frames: the eval-mode tensor-core path reads them directly in its fused stem kernel; every
        # other mode converts + normalises on the device first (one pass, csrc/input.cu)
        if model.training or eval_precision(model) not in ("fp16", "bf16"):
            from .input_pipeline import DeviceInputPipeline

            norm = getattr(model, "rtsds_input_norm", None)
            x = DeviceInputPipeline(None, *(norm if norm is not None else (None, None))).images(x)
    elif x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    if model.training:
        from .bisenet_autograd import bisenet_train_forward

        return bisenet_train_forward(model, x)
    if torch.is_grad_enabled() and x.requires_grad:
        # The reference's eval-mode forward is an ordinary autograd graph.  The eval plan here (folded BatchNorm, fp16, CUDA
        # graph) has no backward; a caller who explicitly asks for input gradients must not get a silently detached tensor.
        raise ops._lib.RtsdsError("BiSeNet eval-mode forward is inference only (no backward through the folded-BatchNorm plan): "
                                  "call it under torch.no_grad(), or use model.train() / rtsds_precision='fp32' training mode "
                                  "for gradients")
    plan = _get_plan(model, x, False)
    with torch.no_grad():          # parameters that require grad do not make the eval output differentiable (documented)
        plan.forward_lowres(x, use_graph=model.rtsds_cuda_graph)
        return plan.logits(plan.z)
