"""CPU stand-ins for the device ops (float64 NCHW tensors play the role of the blocked layout): checks the HOST logic of
the module-level forwards (channel bookkeeping, sizes, crops, head mapping) against the golden vectors."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_numpy as RN
from pbml_mantle_convection_b200 import _lib as L
from pbml_mantle_convection_b200 import ops

_PADMODE = {"zeros": "constant", "constant": "constant", "replicate": "replicate", "reflect": "reflect"}


class Source:
    def __init__(self, t, xform=L.XFORM_NONE, stats=None, gamma=None, beta=None, channels_per_group=4):
        self.t, self.xform, self.gamma, self.beta, self.cpg, self.stats = t, xform, gamma, beta, channels_per_group, stats
        self.inv_count = 1.0 / (channels_per_group * t.shape[-1] * t.shape[-2]) if xform == L.XFORM_GN_GELU else 0.0


def _block_stats(y):
    """(sum, sum^2) per (sample, 4-channel block), float64 -- what the conv epilogues accumulate."""
    B, C, H, W = y.shape
    pad = (-C) % 4
    if pad:
        y = torch.cat([y, y.new_zeros(B, pad, H, W)], 1)
    g = y.reshape(B, -1, 4 * H * W).double()
    return torch.stack([g.sum(-1), (g * g).sum(-1)], -1)


def _apply(src):
    t = src.t
    if src.xform == L.XFORM_GN_GELU and isinstance(src.stats, torch.Tensor) and src.cpg == 4:
        # the device semantics: normalise with the GIVEN sums and count (they may describe more than this tensor: a slab
        # of a decomposed grid carries the GLOBAL statistics)
        C = src.gamma.numel()
        mean = src.stats[..., 0] * src.inv_count
        var = (src.stats[..., 1] * src.inv_count - mean * mean).clamp_min(0.0)
        rstd = torch.rsqrt(var + 1e-5)
        mean_c, rstd_c = mean.repeat_interleave(4, 1)[:, :C], rstd.repeat_interleave(4, 1)[:, :C]
        t = (t[:, :C] - mean_c[:, :, None, None]) * (rstd_c * src.gamma.double()[None])[:, :, None, None] + src.beta.double()[None, :, None, None]
        t = F.gelu(t)
    elif src.xform == L.XFORM_GN_GELU:
        C = src.gamma.numel()
        t = F.group_norm(t[:, :C], C // src.cpg, src.gamma.double(), src.beta.double(), 1e-5)
        t = F.gelu(t)
    elif src.xform == L.XFORM_GELU:
        t = F.gelu(t)
    return t


def install(mp):
    mp.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    mp.setattr(ops, "Source", Source)
    mp.setattr(ops, "pack_nchw", lambda x: x.double())
    mp.setattr(ops, "unpack_nchw", lambda y, c: y[:, :c])
    mp.setattr(ops, "pack_conv_weight", lambda w, ch: w.double())
    mp.setattr(ops, "pack_conv_weight_row", lambda w, ch: w.double())
    mp.setattr(ops, "umma_supported", lambda cout, ch: False)
    mp.setattr(ops, "pad_vec", lambda v, c, dev, fill=0.0: torch.full((c,), fill, dtype=torch.float64) if v is None
               else v.detach().double().reshape(-1))

    def conv_fwd(sources, wpk, bias, cout, k, pad_mode, epi_act=0, want_stats=False, want_chan_sum=False, impl="auto", **kw):
        x = torch.cat([_apply(s) for s in sources], 1)
        x = x[:, :wpk.shape[1]]
        xp = F.pad(x, (k // 2,) * 4, mode=_PADMODE[pad_mode])
        y = F.conv2d(xp, wpk, bias[:cout])
        if epi_act == L.ACT_GELU:
            y = F.gelu(y)
        return y, (_block_stats(y) if want_stats else None), (y.sum(dim=(2, 3)) if want_chan_sum else None)

    mp.setattr(ops, "conv_fwd", conv_fwd)
    mp.setattr(ops, "block_stats", _block_stats)
    mp.setattr(ops, "chan_sums", lambda y: y.double().sum(dim=(2, 3)))
    # learned 9-region conv (two launches on the device): the "filter images" are the full filter banks themselves and the
    # stand-in is the numpy oracle of the reference's nine-region stitch over the concatenated, transformed sources
    mp.setattr(ops, "row_supported", lambda cout, k, ch: True)
    mp.setattr(ops, "pack_edge9_weights", lambda ws, ch: [w.double() for w in ws])

    def conv_learned9(sources, wpk, wrow, wedge, bias, cout, k, epi_act=0, want_stats=False, want_chan_sum=False, impl="auto"):
        x = torch.cat([_apply(s) for s in sources], 1)[:, :wpk.shape[1]]
        sd = {"conv.weight": wpk.numpy(), "learnable_bias": bias[:cout].numpy().reshape(1, cout, 1, 1)}
        for name, w in zip(ops.EDGE9_REGIONS, wedge):
            sd[name + ".weight"] = w.numpy()
        y = torch.tensor(RN.boundary_learned_conv(x.numpy(), sd, "", k, cout))
        if epi_act == L.ACT_GELU:
            y = F.gelu(y)
        csum = y.sum(dim=(2, 3)) if want_chan_sum else None
        return y, ("stats" if want_stats else None), csum

    mp.setattr(ops, "conv_learned9", conv_learned9)
    mp.setattr(ops, "finalize_nchw", lambda src, c: _apply(src)[:, :c])
    mp.setattr(ops, "avgpool2", lambda src: F.avg_pool2d(_apply(src), 2))
    mp.setattr(ops, "bicubic_up", lambda src, H, W, staged=False: F.interpolate(_apply(src), size=(H, W), mode="bicubic",
                                                                                 align_corners=False))

    def head(yb, csum, members, a_bound, kind, p_pred, want_uvmax=True):
        assert kind == L.HEAD_CURL
        y = yb.numpy()
        u, v = RN._curl_uv(y[:, 0] * a_bound)
        if members is not None:  # velocity un-scaling (pbmc_member.scaler)
            sc = members.double().numpy()[:, 6][:, None, None]
            u, v = u * sc, v * sc
        n = y.shape[-1] * y.shape[-2]
        p = torch.tensor(y[:, 1] - (csum[:, 1].numpy() / n)[:, None, None]) if p_pred else None
        uvmax = None
        if want_uvmax:  # max over the interior, as float32 bits (non-negative floats order like their bit patterns)
            m = np.maximum(np.abs(u[:, 1:-1, 1:-1]).max(axis=(1, 2)), np.abs(v[:, 1:-1, 1:-1]).max(axis=(1, 2)))
            uvmax = torch.tensor(m.astype(np.float32)).view(torch.int32)
        return torch.tensor(u), torch.tensor(v), p, uvmax

    mp.setattr(ops, "head", head)
    from pbml_mantle_convection_b200 import pytorch_networks_convae as M

    mp.setattr(M, "_stats_of_blocked", lambda yb, c: None)
    mp.setattr(M, "_require_cuda_device", lambda dev: None)
    mp.setattr(ops, "stencil_coefs", lambda coord64, lo, hi: torch.zeros(3, coord64.numel()))

    def build_input(T, xc, yc, ycc, members, want_V=False):
        # pbmc_member: raq_nd, fkt_nd, fkp_nd, ln fkt, ln fkp, raq, scaler, 0 (ops.member_values)
        m = members.double()
        T4 = T.double()[:, None]
        V = torch.clip(torch.exp(m[:, 3].view(-1, 1, 1, 1) * (0.0 - T4) + m[:, 4].view(-1, 1, 1, 1) * (1.0 - ycc.double())[None, None]),
                       1e-8, 1.0)
        one = torch.ones_like(T4)
        inp = torch.cat([(xc.double() / 4.0)[None, None].expand_as(T4), (yc.double() / 4.0)[None, None].expand_as(T4),
                         torch.log10(V) / 8, one * m[:, 0].view(-1, 1, 1, 1), one * m[:, 1].view(-1, 1, 1, 1),
                         one * m[:, 2].view(-1, 1, 1, 1), T4], 1)
        return inp, (V[:, 0] if want_V else None)

    mp.setattr(ops, "build_input", build_input)
