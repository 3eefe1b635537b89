"""Integer inference engine for the quantized ViT/DeiT forward (reference dataflow: models/vit_fquant.py:334-407,
489-596,830-939; models/layers_quant.py:348-393,462-497 - restated as integer codes in SURVEY.md 8a').

What the reference does per forward in fp32 (re-smooth and re-quantize every weight, 6 elementwise ATen ops per
QAct, ~35 per LayerNorm, ~40 per softmax) is split here into
  * a *plan* (once per bit_config): int8 weight codes with the PoT smoothing folded in, per-column epilogue
    vectors, softmax tables, LayerNorm shift vectors - all resident in HBM;
  * a *program* (once per batch size): the kernel argument blocks over a fixed int8 workspace;
  * the launch sequence, optionally captured into a CUDA graph (7 launches per block, 5 outside):
        patchify(qact_input) -> GEMM[embed epilogue] -> cls rows
        per block: LN1 -> GEMM[qkv, requant] -> attention -> GEMM[proj, residual] -> LN2 -> GEMM[fc1, GELU] -> GEMM[fc2, residual]
        LN(cls rows) -> GEMM[head, dequant]
Only int8 tensors cross HBM between kernels: r/r2 [B*197, D], qkv [B*197, 3D], attention out, MLP hidden.
"""
import os

import torch

from . import intmath, ops
from .ptq import QIntLayerNorm


def _vec(t, n, dev):
    t = torch.as_tensor(t).detach().reshape(-1).to(device=dev, dtype=torch.float32)
    return (t.expand(n) if t.numel() == 1 else t).contiguous()


class EngineNotApplicable(NotImplementedError):
    """the calibrated model uses a configuration the integer program does not cover (Config(ptf=False) / Config(lis=False) /
    non-int8 activation bit types); the model then evaluates module by module (forward_eager), as the reference does"""


def _act(qact, who, symmetric=False):
    """(scale [n] fp32, zero point as a python float) of a calibrated QAct.  Asymmetric observers (omse.py:30-57) give a non-zero
    integer zero point inside [-128,127]; the codes stay int8.  symmetric=True: a channel-wise PTF quantizer (ptf.py:120)."""
    q = qact.quantizer
    if q.scale is None:
        raise RuntimeError("%s is not calibrated (run the calibrate -> model_quant flow or load_quant_state first)" % who)
    if q.bit_type.name != "int8":
        raise EngineNotApplicable("%s: activation bit type %s (engine carries int8 codes)" % (who, q.bit_type.name))
    zp = q.zero_point.detach().reshape(-1).float()
    if symmetric or zp.numel() != 1:
        if bool((zp != 0).any()):
            raise EngineNotApplicable("%s: channel-wise quantizer with zero points" % who)
        return q.scale.detach().reshape(-1).float(), 0.0
    return q.scale.detach().reshape(-1).float(), float(zp)


def _sym_scale(qact, who):
    return _act(qact, who, symmetric=True)[0]


class _Gemm:
    """device-resident pieces of one fused GEMM; in_zp: zero point of the input codes -> zp_corr[n] = in_zp * sum_k W[n,k]"""

    def __init__(self, lin, weight, bit, in_scale, dev, in_zp=0.0):
        lin._set_bits(bit)
        self.W, ws = lin.weight_codes(weight)
        self.N, self.K = self.W.shape
        self.acc_scale = (in_scale.reshape(-1).to(dev) * ws.to(dev)).contiguous()
        self.bias = None if lin.bias is None else lin.bias.detach().float().contiguous()
        self.zp_corr = (self.W.to(torch.int32).sum(dim=1) * int(in_zp)).to(torch.int32).contiguous() if in_zp else None


class VitPlan:
    def __init__(self, model, bits):
        m = model
        dev = m.cls_token.device
        D, L = m.embed_dim, m.depth
        if len(bits) != 4 * L + 2:
            raise ValueError("bit_config needs %d entries (1 + 4*depth + 1), got %d" % (4 * L + 2, len(bits)))
        if any(b not in (4, 8) for b in bits):
            raise ValueError("bit_config entries must be 4 or 8 (registered weight bit types int4/int8)")
        if not all(isinstance(x, QIntLayerNorm) and x.mode == "int" for x in [m.norm] + [b.norm1 for b in m.blocks]):
            raise EngineNotApplicable("the integer engine needs QIntLayerNorm in 'int' mode (Config(ptf=True))")
        if not m.cfg.INT_SOFTMAX:
            raise EngineNotApplicable("the integer engine needs the log-int-softmax (Config(lis=True))")
        self.D, self.L, self.H = D, L, m.num_heads
        self.P = m.patch_size
        self.T = m.patch_embed.num_patches
        fq = lambda v, s, z: ((v / s + z).round().clamp(-128, 127) - z) * s          # quantizer/uniform.py:83-86,125 on plan-time constants
        # ---- stem
        pe = m.patch_embed
        self.input_quant = bool(m.input_quant)
        s_pe, self.z_pe = _act(pe.qact, "patch_embed.qact")
        s_e, self.z_e = _act(m.qact_embed, "qact_embed")
        s_p, z_p = _act(m.qact_pos, "qact_pos")
        s_pe, s_e, s_p = s_pe.to(dev), s_e.to(dev), s_p.to(dev)
        if self.input_quant:
            s_in, self.z_in = _act(m.qact_input, "qact_input")
            self.s_in = float(s_in)
            self.g_embed = _Gemm(pe.proj, pe.proj.weight, bits[0], s_in, dev, self.z_in)
        else:
            # ViT-L (vit_fquant.py:1063): raw fp32 pixels meet fake-quantized weights, so the patch embedding is an
            # fp32 GEMM in the reference too; only its output enters the integer domain (patch_embed.qact).
            pe.proj._set_bits(bits[0])
            codes, ws = pe.proj.weight_codes(pe.proj.weight)
            self.w_embed_hat = (codes.float() * ws.reshape(-1, 1)).contiguous()
            self.b_embed = pe.proj.bias.detach().float().contiguous()
        self.s_pe = _vec(s_pe, 1, dev)
        self.s_pe_f = float(s_pe)
        self.s_e = float(s_e)
        self.s_e_vec = _vec(s_e, 1, dev)
        s0 = _vec(_sym_scale(m.qact1, "qact1"), D, dev)
        pos = m.pos_embed.detach().float()
        self.pos_hat = fq(pos, s_p, z_p).reshape(self.T + 1, D).contiguous()
        cls = m.cls_token.detach().float().reshape(1, D)
        cls_hat = fq(cls, s_e, self.z_e)
        self.cls_row = ((cls_hat + self.pos_hat[0:1]) / s0).round().clamp(-128, 127).to(torch.int8).reshape(D).contiguous()
        self.s_r0 = s0
        # ---- blocks
        self.blocks = []
        last = s0
        for i, blk in enumerate(m.blocks):
            b4 = bits[4 * i + 1: 4 * i + 5]
            a, mlp = blk.attn, blk.mlp
            if a.channel_scale is None or mlp.channel_scale is None:
                raise RuntimeError("block %d is not calibrated" % i)
            p = {}
            cs_a = a.best_scale[[4, 8].index(b4[0])].detach().float().to(dev)
            cs_m = mlp.best_scale[[4, 8].index(b4[2])].detach().float().to(dev)
            (a0, z_a0), (a1, z_a1) = _act(a.qact0, "attn.qact0"), _act(a.qact1, "attn.qact1")
            (as_, z_as), (a2, z_a2) = _act(a.qact_attn1, "attn.qact_attn1"), _act(a.qact2, "attn.qact2")
            (m0, z_m0), (m1, z_m1) = _act(mlp.qact0, "mlp.qact0"), _act(mlp.qact1, "mlp.qact1")
            a0, a1, as_, a2, m0, m1 = (t.to(dev) for t in (a0, a1, as_, a2, m0, m1))
            for sc, nm in ((a0, "attn.qact0"), (a1, "attn.qact1"), (as_, "attn.qact_attn1"), (a2, "attn.qact2"), (m0, "mlp.qact0"), (m1, "mlp.qact1")):
                if sc.numel() != 1:
                    raise EngineNotApplicable("%s must be layer-wise" % nm)
            p["ln1"] = self._ln(blk.norm1, last, a0 * cs_a, cs_a, float(a0), dev, z_a0)
            p["qkv"] = _Gemm(a.qkv, a.qkv.weight * cs_a.reshape(1, -1), b4[0], a0, dev, z_a0)
            p["qkv_out"] = _vec(a1, 3 * D, dev)
            p["qkv_zp"] = z_a1
            p["qkv_pot"] = intmath.is_pot(a1) and intmath.is_pot(p["qkv"].acc_scale) and not (z_a0 or z_a1)
            dh = D // m.num_heads
            p["score_mult"] = float(a1.double() * a1.double() * a.scale / as_.double())
            p["out_mult"] = float(a1.double() / a2.double() / 32768.0)
            p["att_zp"] = (int(z_a1), z_as, z_a2)
            p["lut"] = intmath.lut_to_device(intmath.build_softmax_lut(as_), dev)
            p["proj"] = _Gemm(a.proj, a.proj.weight, b4[1], a2, dev, z_a2)
            p["proj_mid"] = _vec(_sym_scale(a.qact3, "attn.qact3"), D, dev)
            p["res1_scale"] = last
            s_b2 = _vec(_sym_scale(blk.qact2, "block.qact2"), D, dev)
            p["proj_out"] = s_b2
            p["ln2"] = self._ln(blk.norm2, s_b2, m0 * cs_a, cs_m, float(m0), dev, z_m0)   # out grid uses attn's scale (Q7)
            p["fc1"] = _Gemm(mlp.fc1, mlp.fc1.weight * cs_m.reshape(1, -1), b4[2], m0, dev, z_m0)
            p["fc1_out"] = _vec(m1, p["fc1"].N, dev)
            p["fc1_zp"] = z_m1
            p["fc1_pot"] = intmath.is_pot(m1) and not (z_m0 or z_m1)
            # step tables for any output quantizer (thresholds bisected on the reference's own fl(gelu / scale) + zp)
            p["gelu_tab"] = ops.gelu_table(float(m1), dev, zp=float(z_m1))
            p["fc2"] = _Gemm(mlp.fc2, mlp.fc2.weight, b4[3], m1, dev, z_m1)
            p["fc2_mid"] = _vec(_sym_scale(mlp.qact2, "mlp.qact2"), D, dev)
            s_b4 = _vec(_sym_scale(blk.qact4, "block.qact4"), D, dev)
            p["fc2_out"] = s_b4
            p["dh"] = dh
            self.blocks.append(p)
            last = s_b4
        # ---- tail
        q2, z_q2 = _act(m.qact2, "qact2")
        q2 = q2.to(dev)
        ones = torch.ones(D, device=dev)
        self.ln_f = self._ln(m.norm, last, q2, ones, float(q2), dev, z_q2)
        self.head = _Gemm(m.head, m.head.weight, bits[-1], q2, dev, z_q2)
        ao, self.z_ao = _act(m.act_out, "act_out")
        self.head_out = _vec(ao, self.head.N, dev)
        self.head_pot = intmath.is_pot(ao) and intmath.is_pot(self.head.acc_scale) and not (z_q2 or self.z_ao)

    @staticmethod
    def _ln(norm, in_scale, out_scale, post_div, next_scale, dev, next_zp=0.0):
        C = norm.weight.numel()
        in_scale = _vec(in_scale, C, dev)
        s1 = in_scale.min()
        return dict(in_mult=(in_scale / s1).round().contiguous(), s1=float(s1),
                    gamma=norm.weight.detach().float().contiguous(), beta=norm.bias.detach().float().contiguous(),
                    out_scale=_vec(out_scale, C, dev), post_div=_vec(post_div, C, dev), next_scale=next_scale, next_zp=next_zp,
                    pot=(not next_zp) and intmath.is_pot(out_scale) and intmath.is_pot(post_div) and intmath.is_pot(torch.tensor(next_scale)))


class _Lru(dict):
    """insertion-ordered dict with a size bound: the least recently used entry is dropped (plans hold a packed weight set,
    graphs their captured launches; a mixed-precision search visits hundreds of bit_configs)"""

    def __init__(self, cap):
        super().__init__()
        self.cap = cap

    def get_or(self, key, make):
        if key in self:
            v = self.pop(key)
        else:
            v = make()
            while len(self) >= self.cap:
                self.pop(next(iter(self)))
        self[key] = v
        return v


_ATT_AUTOTUNE = os.environ.get("P2VIT_ATT_AUTOTUNE", "1") != "0"


class _AttentionStep:
    """One attention launch of a program.  The tcgen05 kernel has two ways to the log2 probabilities (p2v_attention_args.prob_mode:
    reciprocal + guard band + redo, or the exactly rounded quotient for every score) with identical codes and data-dependent
    speed: coarse score scales put a fifth of ViT-B's 16-score units on exact ties, where the redo path costs more than the
    exact quotient everywhere would.  tune() runs both on the batch in the workspace and keeps the faster one."""

    def __init__(self, args, simt):
        self.args, self.simt, self.tuned = args, simt, bool(simt)

    def __call__(self):
        ops.attention(self.args, simt=self.simt)

    def tune(self):
        best = None
        for mode in (0, 1):
            self.args.prob_mode = mode
            self()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self()
            self()
            e1.record()
            e1.synchronize()
            t = e0.elapsed_time(e1)
            if best is None or t < best[0]:
                best = (t, mode)
        self.args.prob_mode = best[1]
        self.tuned = True
        self()


class VitEngine:
    MAX_PLANS = 8         # packed weight sets kept (one per bit_config, least recently used out first)
    MAX_PROGRAMS = 16     # (bit_config, batch) programs / graphs kept

    def __init__(self, model, use_graph=True, simt_gemm=False):
        self.model = model
        self.use_graph = use_graph
        self.simt_gemm = simt_gemm      # tests only: route GEMMs and attention through the dp4a cross-check kernels
        self.plans, self.programs, self.graphs = _Lru(self.MAX_PLANS), _Lru(self.MAX_PROGRAMS), _Lru(2 * self.MAX_PROGRAMS)
        self.workspaces = _Lru(4)       # one int8 workspace per batch size, shared by every bit_config's program
        self._pixel_luts = {}

    def clear(self):
        """drop every cached plan, program, graph and workspace (e.g. between phases of a search)"""
        for d in (self.graphs, self.programs, self.plans, self.workspaces, self._pixel_luts):
            d.clear()

    def _workspace(self, pl, B):
        dev = self.model.cls_token.device
        D, T = pl.D, pl.T
        R = B * (T + 1)

        def make():
            i8 = lambda *s: torch.empty(s, dtype=torch.int8, device=dev)
            return dict(img=torch.empty((B, 3, T_side(pl), T_side(pl)), dtype=torch.float32, device=dev),
                        cols=i8(B * T, 3 * pl.P * pl.P), ra=i8(R, D), rb=i8(R, D), ln=i8(R, D), qkv=i8(R, 3 * D), ao=i8(R, D),
                        hid=i8(R, pl.blocks[0]["fc1"].N), cls=i8(B, D),
                        logits=torch.empty((B, pl.head.N), dtype=torch.float32, device=dev), logit_codes=i8(B, pl.head.N))
        return self.workspaces.get_or(B, make)

    # ---- argument blocks for one (bit_config, batch) over the batch size's workspace
    def _program(self, bits, B):
        key = (bits, B)
        if key in self.programs:
            return self.programs.get_or(key, None)
        pl = self.plans.get_or(bits, lambda: VitPlan(self.model, list(bits)))
        ws_before = self.workspaces.get(B)
        ws = self._workspace(pl, B)
        if ws_before is not ws:      # the workspace of this batch size was evicted and rebuilt: programs / graphs over the old one go
            for d in (self.programs, self.graphs):
                for k in [k for k in d if k[1] == B]:
                    d.pop(k)
        D, T, H = pl.D, pl.T, pl.H
        R = B * (T + 1)
        steps = []
        if pl.input_quant:
            g = pl.g_embed
            steps.append(("patchify", lambda: ops.quantize_patchify(ws["img"], pl.P, pl.s_in, pl.z_in, out=ws["cols"])))
            steps.append(("embed", self._gemm(ops.gemm_args(ws["cols"], g.W, ops.EPI_EMBED, g.acc_scale, bias=g.bias, out_scale=pl.s_r0,
                                                            mid_scale=pl.s_pe, pos=pl.pos_hat, aux_scale=pl.s_e, tokens_per_image=T,
                                                            out_i8=ws["ra"], zp_corr=g.zp_corr, mid_zp=pl.z_pe, aux_zp=pl.z_e))))
        else:
            # ViT-L: fp32 pixels x dequantized weights on the CUDA cores (csrc/sgemm.cu), patch gather and the whole
            # patch_embed.qact -> qact_embed -> + pos -> qact1 chain fused; the graph holds no library kernel
            steps.append(("patchify", lambda: None))
            steps.append(("embed", lambda: ops.embed_f32(ws["img"], pl.P, pl.w_embed_hat, pl.b_embed, pl.s_pe_f, pl.z_pe, pl.s_e, pl.z_e,
                                                         pl.pos_hat, pl.s_r0, ws["ra"])))
        steps.append(("cls", lambda: ops.fill_cls_rows(ws["ra"], pl.cls_row, B, T, D)))
        for i, p in enumerate(pl.blocks):
            pre = "blocks.%d." % i
            steps.append((pre + "norm1", self._ln(p["ln1"], ws["ra"], R, D, D, ws["ln"])))
            g = p["qkv"]
            steps.append((pre + "attn.qact1", self._gemm(ops.gemm_args(ws["ln"], g.W, ops.EPI_REQUANT, g.acc_scale, bias=g.bias,
                                                                       out_scale=p["qkv_out"], out_i8=ws["qkv"], pot=p["qkv_pot"],
                                                                       zp_corr=g.zp_corr, out_zp=p["qkv_zp"]))))
            at = ops.attention_args(ws["qkv"], ws["ao"], B, T + 1, H, p["dh"], p["score_mult"], p["out_mult"], p["lut"],
                                    zp_qkv=p["att_zp"][0], zp_score=p["att_zp"][1], zp_out=p["att_zp"][2])
            steps.append((pre + "attn.qact2", _AttentionStep(at, self.simt_gemm)))
            g = p["proj"]
            steps.append((pre + "qact2", self._gemm(ops.gemm_args(ws["ao"], g.W, ops.EPI_RESIDUAL, g.acc_scale, bias=g.bias,
                                                                  out_scale=p["proj_out"], mid_scale=p["proj_mid"],
                                                                  res_scale=p["res1_scale"], res=ws["ra"], out_i8=ws["rb"],
                                                                  pot=intmath.is_pot(g.acc_scale), zp_corr=g.zp_corr))))
            steps.append((pre + "norm2", self._ln(p["ln2"], ws["rb"], R, D, D, ws["ln"])))
            g = p["fc1"]
            steps.append((pre + "mlp.qact1", self._gemm(ops.gemm_args(ws["ln"], g.W, ops.EPI_GELU, g.acc_scale, bias=g.bias,
                                                                      out_scale=p["fc1_out"], out_i8=ws["hid"], pot=p["fc1_pot"], gelu_table=p["gelu_tab"],
                                                                      zp_corr=g.zp_corr, out_zp=p["fc1_zp"]))))
            g = p["fc2"]
            steps.append((pre + "qact4", self._gemm(ops.gemm_args(ws["hid"], g.W, ops.EPI_RESIDUAL, g.acc_scale, bias=g.bias,
                                                                  out_scale=p["fc2_out"], mid_scale=p["fc2_mid"],
                                                                  res_scale=p["proj_out"], res=ws["rb"], out_i8=ws["ra"],
                                                                  pot=intmath.is_pot(g.acc_scale), zp_corr=g.zp_corr))))
        steps.append(("qact2", self._ln(pl.ln_f, ws["ra"], B, D, (T + 1) * D, ws["cls"])))
        g = pl.head
        steps.append(("act_out", self._gemm(ops.gemm_args(ws["cls"], g.W, ops.EPI_DEQUANT, g.acc_scale, bias=g.bias, out_scale=pl.head_out,
                                                          out_f32=ws["logits"], out_i8=ws["logit_codes"], pot=pl.head_pot,
                                                          zp_corr=g.zp_corr, out_zp=pl.z_ao))))
        # which workspace tensor holds each step's result (for per-op parity taps)
        outs = {"patchify": "cols", "embed": "ra", "cls": "ra", "qact2": "cls", "act_out": "logits"}
        for i in range(pl.L):
            pre = "blocks.%d." % i
            outs.update({pre + "norm1": "ln", pre + "attn.qact1": "qkv", pre + "attn.qact2": "ao", pre + "qact2": "rb",
                         pre + "norm2": "ln", pre + "mlp.qact1": "hid", pre + "qact4": "ra"})
        prog = dict(ws=ws, steps=steps, outs=outs, plan=pl)
        return self.programs.get_or(key, lambda: prog)

    def _gemm(self, args):
        simt = self.simt_gemm
        return lambda: ops.gemm(args, simt=simt)

    @staticmethod
    def _ln(p, x, rows, C, stride, out):
        a = ops.layernorm_args(x, rows, C, stride, p["in_mult"], p["s1"], p["gamma"], p["beta"], p["out_scale"], p["post_div"],
                               p["next_scale"], p["pot"], out_i8=out, next_zp=p["next_zp"])
        return lambda: ops.layernorm(a)

    def launches_per_forward(self, bit_config):
        return (5 if self.model.input_quant else 4) + 7 * self.model.depth

    def static_input(self, B, bit_config):
        """the fp32 image buffer the program reads; copy a batch into it (e.g. straight from pinned host memory) and call
        run_static() to skip the device-to-device copy of __call__."""
        return self._program(tuple(bit_config), B)["ws"]["img"]

    def run_static(self, B, bit_config, taps=None):
        bits = tuple(bit_config)
        prog = self._program(bits, B)
        if taps is not None:
            for name, fn in prog["steps"]:
                fn()
                taps[name] = prog["ws"][prog["outs"][name]].clone()
            return prog["ws"]["logits"]
        if not self.use_graph:
            for _, fn in prog["steps"]:
                fn()
            return prog["ws"]["logits"]
        self._graph(prog, (bits, B), 0).replay()
        return prog["ws"]["logits"]

    def _graph(self, prog, key, first):
        """CUDA graph of prog's steps[first:] (first = 1: everything after patchify, see __call__)"""
        def capture():
            for _, fn in prog["steps"][first:]:   # eager warm-up: sets kernel attributes, loads modules
                if isinstance(fn, _AttentionStep) and not fn.tuned and _ATT_AUTOTUNE:
                    fn.tune()             # (on the batch in the workspace: the callers below patchify theirs before the first capture)
                else:
                    fn()
            torch.cuda.synchronize()
            g = ops.capture_graph(lambda: [fn() for _, fn in prog["steps"][first:]])
            return g, prog          # the graph replays raw pointers into the program's plan and workspace: it keeps them alive
        return self.graphs.get_or(key, capture)[0]

    def __call__(self, x, bit_config, taps=None):
        if not x.is_cuda:
            raise RuntimeError("p2vit_b200: the quantized forward runs on the GPU only (input is on %s)" % x.device)
        B = x.shape[0]
        bits = tuple(bit_config)
        prog = self._program(bits, B)
        img, pl = prog["ws"]["img"], prog["plan"]
        if x.dtype == torch.uint8:
            return self._call_u8(x, prog, bits, B, taps)
        if (taps is None and self.use_graph and pl.input_quant and x.dtype == torch.float32 and x.is_contiguous()
                and x.shape == img.shape and x.data_ptr() != img.data_ptr()):
            # the only kernel that reads the images is patchify (qact_input + patch gather): launch it on the caller's tensor and
            # replay the graph of the rest - no 4-byte-per-pixel device-to-device copy into the program's own input buffer
            ops.quantize_patchify(x, pl.P, pl.s_in, pl.z_in, out=prog["ws"]["cols"])
            self._graph(prog, (bits, B, "after patchify"), 1).replay()      # (a first use warms up - and picks the attention modes - on this batch)
            return prog["ws"]["logits"].clone()
        if x.data_ptr() != img.data_ptr():
            img.copy_(x)
        return self.run_static(B, bit_config, taps).clone()

    def _call_u8(self, x, prog, bits, B, taps):
        """8-bit pixels: ToTensor + Normalize + qact_input through the per-channel code table (ops.pixel_code_table; the
        normalisation is the one set with model.set_pixel_normalization), then the graph of everything after patchify."""
        pl, ws = prog["plan"], prog["ws"]
        norm = getattr(self.model, "pixel_norm", None)
        if norm is None:
            raise RuntimeError("uint8 input needs model.set_pixel_normalization(mean, std) (test_quant.py:112-127)")
        if not pl.input_quant:
            raise NotImplementedError("uint8 input needs an input quantizer (input_quant=True models)")
        if taps is not None or tuple(x.shape) != tuple(ws["img"].shape) or not x.is_contiguous():
            raise ValueError("uint8 input: contiguous [B,3,%d,%d] batch expected (no taps)" % tuple(ws["img"].shape[2:]))
        key = (norm, float(pl.s_in), float(pl.z_in))
        if key not in self._pixel_luts:
            self._pixel_luts[key] = ops.pixel_code_table(norm[0], norm[1], pl.s_in, x.device, pl.z_in)
        if self.use_graph:
            ops.patchify_u8(x, self._pixel_luts[key], pl.P, out=ws["cols"])
            self._graph(prog, (bits, B, "after patchify"), 1).replay()
        else:
            ops.patchify_u8(x, self._pixel_luts[key], pl.P, out=ws["cols"])
            for _, fn in prog["steps"][1:]:
                fn()
        return ws["logits"].clone()


def T_side(pl):
    return int(round(pl.T ** 0.5)) * pl.P
