"""Integer inference engine for the quantized Swin forward (reference dataflow: models/swin_quant.py:204-254, 397-448,
503-524, 883-914; models/layers_quant.py:348-393, 462-497 - restated as integer codes in SURVEY.md 8a'').

Same split as engine.py: a *plan* (int8 weight codes with the Mlp smoothing folded in, per-column epilogue vectors, softmax
tables, dequantized relative-position bias, shift-mask labels, row maps), a *program* per batch size over a fixed int8
workspace, and the launch sequence captured into a CUDA graph:

    patchify(qact_input) -> GEMM[4x4 conv -> qact_before_norm] -> LN[patch_embed.norm -> patch_embed.qact]
    per block: LN1[-> qact1, stored in shifted-window order] -> GEMM[qkv -> attn.qact1] -> window attention
               [q k^T -> qact_attn1 -> + bias -> qact2 -> (+ mask) -> log2 softmax -> P v -> qact3]
               -> GEMM[proj -> attn.qact4 -> + shortcut -> qact2(PTF), stored back in token order]
               -> LN2[-> qact3 -> / cs -> mlp.qact0] -> GEMM[fc1 -> GELU -> qact1] -> GEMM[fc2 -> qact2(PTF) -> + x -> qact4(PTF)]
    per stage end: LN[2x2 gather as its input row map, 4C -> qact1] -> GEMM[reduction -> qact2(PTF)]
    LN[norm -> qact2] -> average pool + qact3 -> GEMM[head -> act_out]

Window partition, cyclic shift and their inverses never move data on their own: LN1 scatters its rows through a
token->window row map and the proj GEMM's epilogue gathers the shortcut and scatters its result through the inverse map.
"""
import torch

from . import intmath, ops
from .engine import _Gemm, _sym_scale, _vec
from .ptq import QIntLayerNorm
from .swin import window_partition


def _ln_plan(norm, in_scale, out_scale, post_div, next_scale, dev, expand=1):
    C = norm.weight.numel()
    in_scale = in_scale.reshape(-1)
    if expand != 1:
        in_scale = in_scale.unsqueeze(-1).expand(-1, expand).T.reshape(-1)     # layers.py:296-299
    in_scale = _vec(in_scale, C, dev)
    s1 = in_scale.min()
    return dict(in_mult=(in_scale / s1).round().contiguous(), s1=float(s1),
                gamma=norm.weight.detach().float().contiguous(), beta=norm.bias.detach().float().contiguous(),
                out_scale=_vec(out_scale, C, dev), post_div=_vec(post_div, C, dev), next_scale=float(next_scale),
                pot=intmath.is_pot(out_scale) and intmath.is_pot(post_div) and intmath.is_pot(torch.tensor(float(next_scale))))


def _window_maps(B, H, W, ws, shift, dev):
    """token row (b, h, w) -> row in shifted-window order, and the inverse"""
    hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    hs, wsft = (hh - shift) % H, (ww - shift) % W                      # torch.roll(x, -shift): token h lands at (h - shift) mod H
    win = (hs // ws) * (W // ws) + (wsft // ws)
    pos = (hs % ws) * ws + (wsft % ws)
    per_img = (win * ws * ws + pos).reshape(-1)                         # [H*W]
    L = H * W
    to_win = (per_img.unsqueeze(0) + (torch.arange(B) * L).unsqueeze(1)).reshape(-1).to(torch.int32)
    to_tok = torch.empty_like(to_win)
    to_tok[to_win.long()] = torch.arange(B * L, dtype=torch.int32)
    return to_win.to(dev).contiguous(), to_tok.to(dev).contiguous()


def _merge_map(B, H, W, dev):
    """source rows of the 2x2 gather, order x0 (0,0), x1 (1,0), x2 (0,1), x3 (1,1)  (swin_quant.py:514-519)"""
    i, j = torch.meshgrid(torch.arange(H // 2), torch.arange(W // 2), indexing="ij")
    segs = [(2 * i + di) * W + (2 * j + dj) for di, dj in ((0, 0), (1, 0), (0, 1), (1, 1))]
    per_img = torch.stack(segs, dim=-1).reshape(-1, 4)                   # [L/4, 4]
    full = per_img.unsqueeze(0) + (torch.arange(B) * H * W).reshape(-1, 1, 1)
    return full.reshape(-1).to(torch.int32).to(dev).contiguous()


class SwinPlan:
    def __init__(self, model):
        m = model
        dev = next(m.parameters()).device
        norms = [mod for mod in m.modules() if isinstance(mod, QIntLayerNorm)]
        if not all(n.mode == "int" for n in norms):
            raise NotImplementedError("the integer engine needs QIntLayerNorm in 'int' mode (Config(ptf=True))")
        if not m.cfg.INT_SOFTMAX:
            raise NotImplementedError("the integer engine needs the log-int-softmax (Config(lis=True))")
        if not m.input_quant:
            raise NotImplementedError("Swin factories of the reference all use input_quant=True")
        self.P, self.C0, self.grid, self.ws = m.patch_size, m.embed_dim, m.patch_grid[0], m.window_size
        pe = m.patch_embed
        s_in = _sym_scale(m.qact_input, "qact_input")
        self.s_in = float(s_in)
        self.g_embed = _Gemm(pe.proj, pe.proj.weight, 8, s_in, dev)
        s_bn = _sym_scale(pe.qact_before_norm, "patch_embed.qact_before_norm").to(dev)
        s_pe = _sym_scale(pe.qact, "patch_embed.qact").to(dev)
        self.embed_out = _vec(s_bn, self.C0, dev)
        self.embed_pot = intmath.is_pot(s_bn) and intmath.is_pot(self.g_embed.acc_scale)
        ones = lambda C: torch.ones(C, device=dev)
        self.ln_embed = _ln_plan(pe.norm, s_bn, s_pe, ones(self.C0), float(s_pe), dev)
        last = s_pe
        self.stages = []
        for i, layer in enumerate(m.layers):
            C = self.C0 * 2 ** i
            H = self.grid // 2 ** i
            st = dict(C=C, H=H, heads=m.heads[i], blocks=[], merge=None)
            for blk in layer.blocks:
                a, mlp = blk.attn, blk.mlp
                if mlp.channel_scale is None:
                    raise RuntimeError("model is not calibrated")
                p = dict(ws=blk.window_size, shift=blk.shift_size)
                s1 = _sym_scale(blk.qact1, "block.qact1").to(dev)
                sq = _sym_scale(a.qact1, "attn.qact1").to(dev)
                sa1 = _sym_scale(a.qact_attn1, "attn.qact_attn1").to(dev)
                st_ = _sym_scale(a.qact_table, "attn.qact_table").to(dev)
                sa2 = _sym_scale(a.qact2, "attn.qact2").to(dev)
                sa3 = _sym_scale(a.qact3, "attn.qact3").to(dev)
                sa4 = _sym_scale(a.qact4, "attn.qact4").to(dev)
                s3 = _sym_scale(blk.qact3, "block.qact3").to(dev)
                m0 = _sym_scale(mlp.qact0, "mlp.qact0").to(dev)
                m1 = _sym_scale(mlp.qact1, "mlp.qact1").to(dev)
                for s, nm in ((s1, "qact1"), (sq, "attn.qact1"), (sa1, "attn.qact_attn1"), (sa2, "attn.qact2"), (sa3, "attn.qact3"),
                              (sa4, "attn.qact4"), (s3, "qact3"), (m0, "mlp.qact0"), (m1, "mlp.qact1"), (st_, "attn.qact_table")):
                    if s.numel() != 1:
                        raise NotImplementedError("%s must be layer-wise" % nm)
                p["ln1"] = _ln_plan(blk.norm1, last, s1, ones(C), float(s1), dev)
                p["qkv"] = _Gemm(a.qkv, a.qkv.weight, 8, s1, dev)
                p["qkv_out"] = _vec(sq, 3 * C, dev)
                p["qkv_pot"] = intmath.is_pot(sq) and intmath.is_pot(p["qkv"].acc_scale)
                dh = C // st["heads"]
                T = blk.window_size ** 2
                p["T"], p["dh"] = T, dh
                p["score_mult"] = float(sq.double() * sq.double() * a.scale / sa1.double())
                p["s_attn1"], p["s_attn2"] = float(sa1), float(sa2)
                table = a.relative_position_bias_table.detach().float()
                table_codes = (table / st_).round().clamp(-128, 127)
                table_hat = table_codes * st_
                p["bias"] = a.relative_position_bias(table_hat).contiguous()
                # the same table as int8 codes for the tensor-core kernel: bias[h,i,j] == fl(code * s_table)
                p["bias_codes"] = ops.window_bias_codes(a.relative_position_bias(table_codes).to(torch.int8))
                p["bias_scale"] = float(st_)
                p["out_mult"] = float(sq.double() / sa3.double() / 32768.0)
                p["lut"] = intmath.lut_to_device(intmath.build_softmax_lut(sa2), dev)
                p["mask_code"], p["mask_exp"], p["labels"], p["mask_bits"] = 0, 0, None, None
                if blk.shift_size > 0:
                    sf = sa2.float().cpu().reshape(())
                    x0 = torch.floor(-0.6931 / sf)
                    code = torch.round(torch.tensor(-100.0) / sf)
                    if float(-code) - 255.0 < float(32 * -x0):
                        raise NotImplementedError("attn.qact2 scale %g: the -100 shift mask does not reach the clamped tail of int_exp" % float(sf))
                    p["mask_code"] = int(code)
                    p["mask_exp"] = int(torch.floor((1.0 / 0.35815147) / sf ** 2))       # int_polynomial's c_int: exp_int at 32*x0
                    lab = torch.zeros((1, H, H, 1))
                    cnt = 0
                    for hs in (slice(0, -blk.window_size), slice(-blk.window_size, -blk.shift_size), slice(-blk.shift_size, None)):
                        for wsl in (slice(0, -blk.window_size), slice(-blk.window_size, -blk.shift_size), slice(-blk.shift_size, None)):
                            lab[:, hs, wsl, :] = cnt
                            cnt += 1
                    p["labels"] = window_partition(lab, blk.window_size).reshape(-1, T).to(torch.int8).to(dev).contiguous()
                    p["mask_bits"] = ops.window_mask_bits(p["labels"])
                p["proj"] = _Gemm(a.proj, a.proj.weight, 8, sa3, dev)
                p["proj_mid"] = _vec(sa4, C, dev)
                p["res1_scale"] = _vec(last, C, dev)
                s_b2 = _vec(_sym_scale(blk.qact2, "block.qact2"), C, dev)
                p["proj_out"] = s_b2
                cs = mlp.best_scale[1].detach().float().to(dev)                 # 8-bit entry of the smoothing pool
                p["ln2"] = _ln_plan(blk.norm2, s_b2, s3, cs, float(m0), dev)
                p["fc1"] = _Gemm(mlp.fc1, mlp.fc1.weight * cs.reshape(1, -1), 8, m0, dev)
                p["fc1_out"] = _vec(m1, p["fc1"].N, dev)
                p["fc1_pot"] = intmath.is_pot(m1)
                p["gelu_tab"] = ops.gelu_table(float(m1), dev)
                p["fc2"] = _Gemm(mlp.fc2, mlp.fc2.weight, 8, m1, dev)
                p["fc2_mid"] = _vec(_sym_scale(mlp.qact2, "mlp.qact2"), C, dev)
                s_b4 = _vec(_sym_scale(blk.qact4, "block.qact4"), C, dev)
                p["fc2_out"] = s_b4
                st["blocks"].append(p)
                last = s_b4
            if layer.downsample is not None:
                ds = layer.downsample
                d1 = _sym_scale(ds.qact1, "downsample.qact1").to(dev)
                mg = dict(ln=_ln_plan(ds.norm, last, d1, ones(4 * C), float(d1), dev, expand=4))
                mg["red"] = _Gemm(ds.reduction, ds.reduction.weight, 8, d1, dev)
                mg["out"] = _vec(_sym_scale(ds.qact2, "downsample.qact2"), 2 * C, dev)
                st["merge"] = mg
                last = mg["out"]
            self.stages.append(st)
        Cf = m.num_features
        q2 = _sym_scale(m.qact2, "qact2").to(dev)
        q3 = _sym_scale(m.qact3, "qact3").to(dev)
        self.ln_f = _ln_plan(m.norm, last, q2, ones(Cf), float(q2), dev)
        self.s_q2, self.s_q3 = float(q2), float(q3)
        self.head = _Gemm(m.head, m.head.weight, 8, q3, dev)
        ao = _sym_scale(m.act_out, "act_out")
        self.head_out = _vec(ao, self.head.N, dev)
        self.head_pot = intmath.is_pot(ao) and intmath.is_pot(self.head.acc_scale)
        self.Cf = Cf


class SwinEngine:
    def __init__(self, model, use_graph=True, simt=False):
        self.model, self.use_graph = model, use_graph
        self.simt = simt                # tests only: window attention on the dp4a cross-check kernel
        self.plan, self.programs, self.graphs = None, {}, {}
        self._pixel_luts = {}

    def _program(self, B):
        if B in self.programs:
            return self.programs[B]
        if self.plan is None:
            self.plan = SwinPlan(self.model)
        pl = self.plan
        dev = next(self.model.parameters()).device
        side = pl.grid * pl.P
        L0 = pl.grid * pl.grid
        n0 = B * L0 * pl.C0                                                # bytes of one [rows, C] activation (constant over stages / 2)
        flat = lambda n: torch.empty(n, dtype=torch.int8, device=dev)
        ws = dict(img=torch.empty((B, 3, side, side), dtype=torch.float32, device=dev), cols=flat(B * L0 * 3 * pl.P * pl.P),
                  ra=flat(n0), rb=flat(n0), ln=flat(n0), ao=flat(n0), qkv=flat(3 * n0), hid=flat(4 * n0),
                  pool=flat(B * pl.Cf), logits=torch.empty((B, pl.head.N), dtype=torch.float32, device=dev), logit_codes=flat(B * pl.head.N))
        view = lambda name, rows, C: ws[name][: rows * C].view(rows, C)
        steps = []
        gemm = lambda args: (lambda: ops.gemm(args))

        def ln(p, x, rows, C, out, row_map=None, clamp_mid=False, gather=None):
            # gather: patch merging - the row is the concatenation of four source rows of C / 4 channels (row stride = source row)
            a = ops.layernorm_args(x, rows, C, C // 4 if gather is not None else C, p["in_mult"], p["s1"], p["gamma"], p["beta"], p["out_scale"],
                                   p["post_div"], p["next_scale"], p["pot"], out_i8=out, out_row_map=row_map, clamp_mid=clamp_mid,
                                   in_gather=gather, gather_segs=4 if gather is not None else 0)
            return lambda: ops.layernorm(a)

        R = B * L0
        g = pl.g_embed
        cols = view("cols", R, 3 * pl.P * pl.P)
        steps.append(("patchify", lambda: ops.quantize_patchify(ws["img"], pl.P, pl.s_in, out=cols)))
        steps.append(("patch_embed.qact_before_norm", gemm(ops.gemm_args(cols, g.W, ops.EPI_REQUANT, g.acc_scale, bias=g.bias, out_scale=pl.embed_out,
                                                                         out_i8=view("rb", R, pl.C0), pot=pl.embed_pot))))
        steps.append(("patch_embed.qact", ln(pl.ln_embed, view("rb", R, pl.C0), R, pl.C0, view("ra", R, pl.C0))))
        outs = {"patchify": ("cols", R, 3 * pl.P * pl.P), "patch_embed.qact_before_norm": ("rb", R, pl.C0), "patch_embed.qact": ("ra", R, pl.C0)}
        keep = []                                                          # row maps must outlive the captured graph
        for i, st in enumerate(pl.stages):
            C, H = st["C"], st["H"]
            R = B * H * H
            ra, rb, lnb, ao = (view(n, R, C) for n in ("ra", "rb", "ln", "ao"))
            qkv, hid = view("qkv", R, 3 * C), view("hid", R, 4 * C)
            maps = {}
            for j, p in enumerate(st["blocks"]):
                pre = "layers.%d.blocks.%d." % (i, j)
                key = (p["ws"], p["shift"])
                if key not in maps:
                    maps[key] = _window_maps(B, H, H, p["ws"], p["shift"], dev)
                    keep.append(maps[key])
                to_win, to_tok = maps[key]
                steps.append((pre + "qact1", ln(p["ln1"], ra, R, C, lnb, row_map=to_win)))
                gq = p["qkv"]
                steps.append((pre + "attn.qact1", gemm(ops.gemm_args(lnb, gq.W, ops.EPI_REQUANT, gq.acc_scale, bias=gq.bias, out_scale=p["qkv_out"],
                                                                     out_i8=qkv, pot=p["qkv_pot"]))))
                T = p["T"]
                wa = ops.window_attention_args(qkv, ao, R // T, T, st["heads"], p["dh"], (H // p["ws"]) ** 2, p["score_mult"], p["s_attn1"],
                                               p["s_attn2"], p["bias"], p["labels"], p["mask_code"], p["mask_exp"], p["out_mult"], p["lut"],
                                               out_row_map=to_tok,      # window_reverse + roll in the store: `ao` is in token order
                                               bias_codes=p["bias_codes"], bias_scale=p["bias_scale"], mask_bits=p["mask_bits"])
                steps.append((pre + "attn.qact3", (lambda wa=wa, simt=self.simt: ops.window_attention(wa, simt=simt))))
                gp = p["proj"]
                steps.append((pre + "qact2", gemm(ops.gemm_args(ao, gp.W, ops.EPI_RESIDUAL, gp.acc_scale, bias=gp.bias, out_scale=p["proj_out"],
                                                                mid_scale=p["proj_mid"], res_scale=p["res1_scale"], res=ra, out_i8=rb,
                                                                pot=intmath.is_pot(gp.acc_scale)))))
                steps.append((pre + "mlp.qact0", ln(p["ln2"], rb, R, C, lnb, clamp_mid=True)))
                g1 = p["fc1"]
                steps.append((pre + "mlp.qact1", gemm(ops.gemm_args(lnb, g1.W, ops.EPI_GELU, g1.acc_scale, bias=g1.bias, out_scale=p["fc1_out"],
                                                                    out_i8=hid, pot=p["fc1_pot"], gelu_table=p["gelu_tab"]))))
                g2 = p["fc2"]
                steps.append((pre + "qact4", gemm(ops.gemm_args(hid, g2.W, ops.EPI_RESIDUAL, g2.acc_scale, bias=g2.bias, out_scale=p["fc2_out"],
                                                                mid_scale=p["fc2_mid"], res_scale=p["proj_out"], res=rb, out_i8=ra,
                                                                pot=intmath.is_pot(g2.acc_scale)))))
                outs.update({pre + "qact1": ("ln", R, C), pre + "attn.qact1": ("qkv", R, 3 * C), pre + "attn.qact3": ("ao", R, C),
                             pre + "qact2": ("rb", R, C), pre + "mlp.qact0": ("ln", R, C), pre + "mlp.qact1": ("hid", R, 4 * C),
                             pre + "qact4": ("ra", R, C)})
            if st["merge"] is not None:
                mg = st["merge"]
                pre = "layers.%d.downsample." % i
                src = _merge_map(B, H, H, dev)
                keep.append(src)
                R4 = R // 4
                ln4 = view("ln", R4, 4 * C)
                # the 2x2 neighbourhood gather (swin_quant.py:512-519) is the merge LayerNorm's input row map: no gather kernel, no [R/4, 4C] copy
                steps.append((pre + "qact1", ln(mg["ln"], ra, R4, 4 * C, ln4, gather=src)))
                gr = mg["red"]
                steps.append((pre + "qact2", gemm(ops.gemm_args(ln4, gr.W, ops.EPI_REQUANT, gr.acc_scale, bias=None, out_scale=mg["out"],
                                                                out_i8=view("ra", R4, 2 * C)))))
                outs.update({pre + "qact1": ("ln", R4, 4 * C), pre + "qact2": ("ra", R4, 2 * C)})
        Cf, Tl = pl.Cf, pl.stages[-1]["H"] ** 2
        Rl = B * Tl
        steps.append(("qact2", ln(pl.ln_f, view("ra", Rl, Cf), Rl, Cf, view("ln", Rl, Cf))))
        pool = ws["pool"].view(B, Cf)
        steps.append(("qact3", lambda: ops.avgpool_quant(view("ln", Rl, Cf), pool, B, Tl, Cf, pl.s_q2, pl.s_q3)))
        gh = pl.head
        steps.append(("act_out", gemm(ops.gemm_args(pool, gh.W, ops.EPI_DEQUANT, gh.acc_scale, bias=gh.bias, out_scale=pl.head_out,
                                                    out_f32=ws["logits"], out_i8=ws["logit_codes"].view(B, gh.N), pot=pl.head_pot))))
        outs.update({"qact2": ("ln", Rl, Cf), "qact3": ("pool", B, Cf)})
        prog = dict(ws=ws, steps=steps, outs=outs, keep=keep, view=view)
        self.programs[B] = prog
        return prog

    def launches_per_forward(self):
        prog = next(iter(self.programs.values()))
        return len(prog["steps"])

    def static_input(self, B):
        return self._program(B)["ws"]["img"]

    def run_static(self, B, taps=None):
        prog = self._program(B)
        if taps is not None:
            for name, fn in prog["steps"]:
                fn()
                if name in prog["outs"]:
                    taps[name] = prog["view"](*prog["outs"][name]).clone()
            return prog["ws"]["logits"]
        if not self.use_graph:
            for _, fn in prog["steps"]:
                fn()
            return prog["ws"]["logits"]
        if B not in self.graphs:
            for _, fn in prog["steps"]:
                fn()
            torch.cuda.synchronize()
            self.graphs[B] = ops.capture_graph(lambda: [fn() for _, fn in prog["steps"]])
        self.graphs[B].replay()
        return prog["ws"]["logits"]

    def __call__(self, x, taps=None):
        if not x.is_cuda:
            raise RuntimeError("p2vit_b200: the quantized forward runs on the GPU only (input is on %s)" % x.device)
        B = x.shape[0]
        img = self.static_input(B)
        if x.dtype == torch.uint8:
            return self._call_u8(x, B, taps)
        if x.data_ptr() != img.data_ptr():
            img.copy_(x)
        return self.run_static(B, taps).clone()

    def _call_u8(self, x, B, taps):
        """8-bit pixels through the per-channel code table (see VitEngine._call_u8); steps[0] is the fp32 patchify it replaces."""
        prog, pl = self._program(B), self.plan
        ws = prog["ws"]
        norm = getattr(self.model, "pixel_norm", None)
        if norm is None:
            raise RuntimeError("uint8 input needs model.set_pixel_normalization(mean, std) (test_quant.py:112-127)")
        if taps is not None or tuple(x.shape) != tuple(ws["img"].shape) or not x.is_contiguous():
            raise ValueError("uint8 input: contiguous [B,3,%d,%d] batch expected (no taps)" % tuple(ws["img"].shape[2:]))
        key = (norm, float(pl.s_in))
        if key not in self._pixel_luts:
            self._pixel_luts[key] = ops.pixel_code_table(norm[0], norm[1], pl.s_in, x.device)
        assert prog["steps"][0][0] == "patchify"
        gkey = (B, "after patchify")
        if self.use_graph and gkey not in self.graphs:
            for _, fn in prog["steps"]:
                fn()
            torch.cuda.synchronize()
            self.graphs[gkey] = ops.capture_graph(lambda: [fn() for _, fn in prog["steps"][1:]])
        ops.patchify_u8(x, self._pixel_luts[key], pl.P, out=ws["cols"])
        if self.use_graph:
            self.graphs[gkey].replay()
        else:
            for _, fn in prog["steps"][1:]:
                fn()
        return ws["logits"].clone()
