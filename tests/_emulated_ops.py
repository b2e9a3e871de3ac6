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
        self.t, self.xform, self.gamma, self.beta, self.cpg = t, xform, gamma, beta, channels_per_group


def _apply(src):
    t = src.t
    if src.xform == L.XFORM_GN_GELU:
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
    mp.setattr(ops, "pad_vec", lambda v, c, dev, fill=0.0: torch.full((c,), fill, dtype=torch.float64) if v is None
               else v.detach().double().reshape(-1))

    def conv_fwd(sources, wpk, bias, cout, k, pad_mode, epi_act=0, want_stats=False, want_chan_sum=False, impl="auto", **kw):
        x = torch.cat([_apply(s) for s in sources], 1)
        x = x[:, :wpk.shape[1]]
        xp = F.pad(x, (k // 2,) * 4, mode=_PADMODE[pad_mode])
        return F.conv2d(xp, wpk, bias[:cout]), ("stats" if want_stats else None), None

    mp.setattr(ops, "conv_fwd", conv_fwd)
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
        assert kind == L.HEAD_CURL and members is None
        y = yb.numpy()
        u, v = RN._curl_uv(y[:, 0] * a_bound)
        n = y.shape[-1] * y.shape[-2]
        p = torch.tensor(y[:, 1] - (csum[:, 1].numpy() / n)[:, None, None]) if p_pred else None
        return torch.tensor(u), torch.tensor(v), p, None

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
