"""TEST INFRASTRUCTURE ONLY - CPU oracle for the quantized Swin path of P2-ViT (models/swin_quant.py), built on the
pieces of oracle/port.py (fake_quant, int_layernorm, int_softmax_log2).  Only `tests/` and bench.py's CPU arm import it.

Pinning: tests/golden/swin_micro_minmax.npz comes from the reference's own classes run through oracle/gen_golden_swin.py
(with the four call-site adapters of SURVEY Q4, which change no arithmetic); tests/test_oracle_golden.py checks this port
against its calibrated state, logits and per-module checksums.

Reference map (file:line relative to /root/reference):
  window_partition / window_reverse      models/swin_quant.py:26-59
  WindowAttention.forward                models/swin_quant.py:204-254   (q*scale BEFORE the matmul, quantized relative
                                         position bias table, mask added after qact2)
  SwinTransformerBlock.forward           models/swin_quant.py:397-448   (roll by -3 on odd blocks, SW-MSA mask :365-395)
  PatchMerging.forward                   models/swin_quant.py:503-524   (2x2 gather, LN with in_scale_expand=4)
  SwinTransformer.forward_features       models/swin_quant.py:883-914
  PatchEmbed.forward (with norm)         models/layers_quant.py:462-497
  Mlp.forward (PoT smoothing)            models/layers_quant.py:348-393

The state is always loaded (reference-calibrated golden state or the state exported by the B200 model); this port has no
calibration pass of its own.  `exact_sums=True`: canonical variant as in oracle/port.py (exact row sums, IEEE sqrt, exact
integer accumulation in the matmuls with the head scale factored out of q k^T).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .port import BITS, _ActQ, _WeightQ, fake_quant, int_layernorm, int_softmax_log2, tdiv  # noqa: F401


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def relative_position_index(ws):
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shifted_window_mask(H, W, ws, shift):
    """[nW, ws*ws, ws*ws] of 0 / -100 (swin_quant.py:365-395)"""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


class SwinOracle:
    def __init__(self, sd, embed_dim, depths, num_heads, method="minmax", exact_sums=False, window=7, patch=4, img=224, device="cpu", **_):
        self.dev = torch.device(device)      # "cuda": same restatement on torch's CUDA backend (see oracle/port.py)
        self.sd = {k: v.float().to(self.dev) for k, v in sd.items()}
        self.C0, self.depths, self.heads, self.ws, self.P = embed_dim, tuple(depths), tuple(num_heads), window, patch
        self.grid = img // patch
        self.exact = exact_sums
        A = lambda: _ActQ("int8", "layer_wise", method)
        LN = lambda: _ActQ("int8", "channel_wise", "ptf")
        W = lambda kind="linear_weight": _WeightQ(kind)
        q = self.q = {"qact_input": A(), "patch_embed.proj": W("conv_weight"), "patch_embed.qact_before_norm": A(), "patch_embed.qact": A(),
                      "qact2": A(), "qact3": A(), "head": W(), "act_out": A()}
        for i, depth in enumerate(self.depths):
            for j in range(depth):
                p = "layers.%d.blocks.%d." % (i, j)
                for nm in ("qact1", "attn.qact1", "attn.qact_attn1", "attn.qact_table", "attn.qact2", "attn.qact3", "attn.qact4", "qact3",
                           "mlp.qact0", "mlp.qact1"):
                    q[p + nm] = A()
                for nm in ("qact2", "mlp.qact2", "qact4"):
                    q[p + nm] = LN()
                for nm in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                    q[p + nm] = W()
            if i < len(self.depths) - 1:
                p = "layers.%d.downsample." % i
                q[p + "qact1"], q[p + "reduction"], q[p + "qact2"] = A(), W(), LN()
        self.cs = {}

    def load_state(self, st):
        for nm, qq in self.q.items():
            if isinstance(qq, _ActQ):
                qq.scale = torch.as_tensor(st[nm + ".scale"]).float().to(self.dev)
                qq.zp = torch.as_tensor(st[nm + ".zero_point"]).long().to(self.dev)
            else:
                for bit in ("uint3", "uint4", "int4", "int8"):
                    k = "%s.scale.%s" % (nm, bit)
                    if k in st:
                        qq.scale[bit] = torch.as_tensor(st[k]).float().to(self.dev)
                        qq.zp[bit] = torch.as_tensor(st["%s.zero_point.%s" % (nm, bit)]).long().to(self.dev)
        for k, v in st.items():
            if k.endswith(".channel_scale"):
                self.cs[k[: -len(".channel_scale")]] = torch.as_tensor(v).float().to(self.dev)

    # ---- matmuls: reference fp32 form, or exact integer accumulation (canonical)
    def _lin(self, h, aq, wq, w, bias):
        wm = wq.fq(w, "int8").reshape(w.shape[0], -1)
        if not self.exact or aq.scale.numel() != 1 or bool((aq.zp != 0).any()):
            return F.linear(h, wm, bias)
        ws = wq.scale["int8"].reshape(-1, 1).float()
        acc = (torch.round(h / aq.scale.reshape(-1)).double() @ torch.round(wm / ws).double().T).float()
        y = acc * (aq.scale.reshape(-1).float() * ws.reshape(-1))
        return y if bias is None else y + bias

    def _ln(self, x, name, in_q, out_q, expand=1):
        return int_layernorm(x, in_q.scale, out_q.scale, self.sd[name + ".weight"], self.sd[name + ".bias"], self.exact, expand)

    def _attention(self, x, p, heads, mask, in_q, tap):
        """x: [B_, N, C] dequantized block.qact1 output in window order"""
        q, sd = self.q, self.sd
        B_, N, C = x.shape
        dh = C // heads
        scale = dh ** -0.5
        h = q[p + "attn.qact1"](self._lin(x, in_q, q[p + "attn.qkv"], sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]))
        tap(p + "attn.qact1", h)
        qkv = h.reshape(B_, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
        qh, kh, vh = qkv[0], qkv[1], qkv[2]
        q1, qa1, qa2, qa3 = q[p + "attn.qact1"], q[p + "attn.qact_attn1"], q[p + "attn.qact2"], q[p + "attn.qact3"]
        table = q[p + "attn.qact_table"](sd[p + "attn.relative_position_bias_table"])
        bias = table[self.rpi.view(-1).to(self.dev)].view(N, N, -1).permute(2, 0, 1).contiguous()
        if self.exact:
            cq, ck, cv = (torch.round(t / q1.scale.reshape(())).double() for t in (qh, kh, vh))
            S = (cq @ ck.transpose(-2, -1)).float()
            mult = (q1.scale.reshape(()).double() ** 2 * scale / qa1.scale.reshape(()).double()).float()
            a = torch.clamp(torch.round(S * mult), -128, 127) * qa1.scale.reshape(())
        else:
            a = qa1((qh * scale) @ kh.transpose(-2, -1))
        tap(p + "attn.qact_attn1", a)
        a = qa2(a + bias.unsqueeze(0))
        tap(p + "attn.qact2", a)
        codes = torch.round(a / qa2.scale.reshape(()))
        if mask is not None:
            mask = mask.to(self.dev)
            nW = mask.shape[0]
            a = (a.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
            codes = (codes.view(B_ // nW, nW, heads, N, N) + torch.round(mask / qa2.scale.reshape(())).unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
        pr = int_softmax_log2(a, qa2.scale, 4, self.exact, codes=codes if self.exact else None)
        tap(p + "attn.log_int_softmax", pr)
        if self.exact:
            O = ((pr.double() * 32768.0) @ cv).float()
            omult = (q1.scale.reshape(()).double() / qa3.scale.reshape(()).double() / 32768.0).float()
            h = (torch.clamp(torch.round(O * omult), -128, 127) * qa3.scale.reshape(())).transpose(1, 2).reshape(B_, N, C)
        else:
            h = qa3((pr @ vh).transpose(1, 2).reshape(B_, N, C))
        tap(p + "attn.qact3", h)
        h = q[p + "attn.qact4"](self._lin(h, qa3, q[p + "attn.proj"], sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]))
        tap(p + "attn.qact4", h)
        return h

    @torch.no_grad()
    def forward_quant(self, x, taps=None):
        q, sd, ws = self.q, self.sd, self.ws
        tap = (lambda n, v: taps.__setitem__(n, v.clone())) if taps is not None else (lambda n, v: None)
        B = x.shape[0]
        x = q["qact_input"](x.to(self.dev))
        tap("qact_input", x)
        if self.exact:
            P, gs = self.P, self.grid
            rows = x.reshape(B, 3, gs, P, gs, P).permute(0, 2, 4, 1, 3, 5).reshape(B, gs * gs, 3 * P * P)
            x = self._lin(rows, q["qact_input"], q["patch_embed.proj"], sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"])
        else:
            w = q["patch_embed.proj"].fq(sd["patch_embed.proj.weight"], "int8")
            x = F.conv2d(x, w, sd["patch_embed.proj.bias"], (self.P, self.P)).flatten(2).transpose(1, 2)
        x = q["patch_embed.qact_before_norm"](x)
        tap("patch_embed.qact_before_norm", x)
        x = q["patch_embed.qact"](self._ln(x, "patch_embed.norm", q["patch_embed.qact_before_norm"], q["patch_embed.qact"]))
        tap("patch_embed.qact", x)
        last = q["patch_embed.qact"]
        for i, depth in enumerate(self.depths):
            H = W = self.grid // 2 ** i
            C = self.C0 * 2 ** i
            win = min(ws, H)
            self.rpi = relative_position_index(win)
            for j in range(depth):
                p = "layers.%d.blocks.%d." % (i, j)
                shift = 0 if (j % 2 == 0 or H <= ws) else ws // 2
                shortcut = x
                h = q[p + "qact1"](self._ln(x, p + "norm1", last, q[p + "qact1"]))
                tap(p + "qact1", h)
                h = h.view(B, H, W, C)
                if shift:
                    h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
                hw = window_partition(h, win).view(-1, win * win, C)
                hw = self._attention(hw, p, self.heads[i], shifted_window_mask(H, W, win, shift) if shift else None, q[p + "qact1"], tap)
                h = window_reverse(hw.view(-1, win, win, C), win, H, W)
                if shift:
                    h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
                x = q[p + "qact2"](shortcut + h.view(B, H * W, C))
                tap(p + "qact2", x)
                h = q[p + "qact3"](self._ln(x, p + "norm2", q[p + "qact2"], q[p + "qact3"]))
                tap(p + "qact3", h)
                cs = self.cs[p + "mlp"]
                h = q[p + "mlp.qact0"](h / cs.reshape(1, 1, -1))
                tap(p + "mlp.qact0", h)
                h = q[p + "mlp.qact1"](F.gelu(self._lin(h, q[p + "mlp.qact0"], q[p + "mlp.fc1"], sd[p + "mlp.fc1.weight"] * cs.reshape(1, -1),
                                                        sd[p + "mlp.fc1.bias"])))
                tap(p + "mlp.qact1", h)
                h = q[p + "mlp.qact2"](self._lin(h, q[p + "mlp.qact1"], q[p + "mlp.fc2"], sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]))
                tap(p + "mlp.qact2", h)
                x = q[p + "qact4"](x + h)
                tap(p + "qact4", x)
                last = q[p + "qact4"]
            if i < len(self.depths) - 1:
                p = "layers.%d.downsample." % i
                x = x.view(B, H, W, C)
                x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1).view(B, -1, 4 * C)
                x = q[p + "qact1"](self._ln(x, p + "norm", last, q[p + "qact1"], 4))
                tap(p + "qact1", x)
                x = q[p + "qact2"](self._lin(x, q[p + "qact1"], q[p + "reduction"], sd[p + "reduction.weight"], sd[p + "reduction.bias"]))
                tap(p + "qact2", x)
                last = q[p + "qact2"]
        x = q["qact2"](self._ln(x, "norm", last, q["qact2"]))
        tap("qact2", x)
        if self.exact:   # mean of the codes: exact integer sum, one fp32 product and one fp32 division
            s = q["qact2"].scale.reshape(())
            x = tdiv(torch.round(x / s).double().sum(1).float() * s, x.shape[1])
        else:
            x = F.adaptive_avg_pool1d(x.transpose(1, 2), 1).flatten(1)
        x = q["qact3"](x)
        tap("qact3", x)
        return q["act_out"](self._lin(x, q["qact3"], q["head"], sd["head.weight"], sd["head.bias"]))


def tap_checksums(taps):
    """name -> (sum, sum of squares) in float64: compact fingerprints of the per-module outputs stored in the golden file"""
    return {k: np.array([v.double().sum().item(), (v.double() ** 2).sum().item()]) for k, v in taps.items()}
