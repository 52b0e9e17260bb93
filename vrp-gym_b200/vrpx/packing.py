"""Host-side packing of policy parameters for the CUDA kernels.

Encoder: the kernels read the torch parameters in place (struct vrpx_encoder_weights = borrowed pointers).

Decoder: the reference recomputes, at every decode step, projections that do not depend on the step
(agents/graph_decoder.py:75-94: K/V/`_kp` of every node, graph mean).  Here the chain of shared-weight
linear maps is folded ONCE per parameter version into two dense matrices (in float64, rounded once to f32):

    q~ = A · ctx_parts        per-head query with the key projection folded in:   scores = q~_h · h_n
    q^ = M · c + m_c          V-proj, out_proj, _att_output and _kp folded:        logits = 10 tanh(q^ · h_n)

with c_h = sum_n softmax_n(scores)_hn h_n.  See include/vrpx.h (vrpx_decoder_weights) for every array.
The folding is written with differentiable torch ops so that gradients w.r.t. the packed arrays can be
pulled back to the module parameters with one autograd call (training path).
"""
from __future__ import annotations

import math

import torch

import vrpx

H, E, D = 8, 128, 384
DH = D // H  # 48


def fold_decoder(dec, irp: bool, dtype=torch.float64):
    """Return the dict of packed decoder arrays (torch tensors of `dtype` on the parameters' device).

    dec: GraphDecoder module.  Formulas: SURVEY App. A.3; graph_decoder.py:88-98."""
    att = dec.attention
    Wq = att.q_proj_weight.to(dtype)                  # (384, 384)
    Wk = att.k_proj_weight.to(dtype)                  # (384, 128)
    Wv = att.v_proj_weight.to(dtype)                  # (384, 128)
    bq, _bk, bv = att.in_proj_bias.to(dtype).split(D)  # b_k cancels in the softmax
    Wo, bo = att.out_proj.weight.to(dtype), att.out_proj.bias.to(dtype)
    Wao = dec._att_output.weight.to(dtype)            # (128, 384)
    Wkp = dec._kp.weight.to(dtype)                    # (128, 128)
    f0 = dec._first_node.to(dtype).reshape(E)
    l0 = dec._last_node.to(dtype).reshape(E)

    Wk_h = Wk.view(H, DH, E)

    def kfold(X):  # X (384, m) -> (1024, m): per head W_k,h^T · X[head rows] / sqrt(48)
        m = X.shape[1]
        return torch.einsum("hjd,hjm->hdm", Wk_h, X.reshape(H, DH, m)).reshape(H * E, m) / math.sqrt(DH)

    if irp:
        Wqc = Wq @ dec._context_proj.weight.to(dtype)  # (384, 257): q = W_q · W_ctx · [g, last, load] + b_q
        Wg, Wl, wload = Wqc[:, :E], Wqc[:, E:2 * E], Wqc[:, 2 * E:]
        out = {
            "ag_t": kfold(Wg).T.contiguous(),
            "af_t": None,
            "al_t": kfold(Wl).T.contiguous(),
            "a_q0": kfold((Wl @ l0)[:, None])[:, 0].contiguous(),
            "a_load": kfold(wload)[:, 0].contiguous(),
        }
    else:
        Wg, Wf, Wl = Wq[:, :E], Wq[:, E:2 * E], Wq[:, 2 * E:]
        out = {
            "ag_t": kfold(Wg).T.contiguous(),
            "af_t": kfold(Wf).T.contiguous(),
            "al_t": kfold(Wl).T.contiguous(),
            "a_q0": kfold((Wf @ f0 + Wl @ l0)[:, None])[:, 0].contiguous(),
            "a_load": None,
        }
    out["a_c"] = kfold(bq[:, None])[:, 0].contiguous()
    # rank-48 factors of al_t per head (score-table mode of vrpx_rollout): q' = W_l h / sqrt(48), k' = W_k h
    out["qk_w"] = torch.cat([Wl / math.sqrt(DH), Wk], dim=0).contiguous()   # (768, 128)
    F = (Wkp.T @ Wao @ Wo) / math.sqrt(E)              # (128, 384)
    # m_t[h*128 + d, e] = sum_j F[e, 48h+j] * Wv[48h+j, d]
    out["m_t"] = torch.einsum("ehj,hjd->hde", F.view(E, H, DH), Wv.view(H, DH, E)).reshape(H * E, E).contiguous()
    out["m_c"] = ((Wkp.T @ Wao @ (Wo @ bv + bo)) / math.sqrt(E)).contiguous()
    return out


class PackedDecoder:
    """f32 device copies of the folded arrays + the ctypes struct, cached per parameter version."""

    def __init__(self):
        self._key = None
        self.tensors = None
        self.struct = None

    def __deepcopy__(self, memo):
        return PackedDecoder()  # the cache holds raw device pointers: never copy it, rebuild on demand

    def get(self, dec, irp: bool, device) -> "vrpx.DecoderWeights":
        params = list(dec.parameters())
        key = (irp, str(device)) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            with torch.no_grad():
                folded = fold_decoder(dec, irp)
            self.tensors = {k: (None if v is None else v.to(device=device, dtype=torch.float32).contiguous())
                            for k, v in folded.items()}
            s = vrpx.DecoderWeights()
            for k, v in self.tensors.items():
                setattr(s, k, None if v is None else v.data_ptr())
            self.struct = s
            self._key = key
        return self.struct


def encoder_struct(enc, device) -> "vrpx.EncoderWeights":
    """Borrow the encoder's parameters / BatchNorm buffers (they must be contiguous f32 on `device`)."""
    def p(t):
        assert t.device == device and t.dtype == torch.float32 and t.is_contiguous(), "encoder params must be f32 CUDA"
        return t.data_ptr()

    w = vrpx.EncoderWeights()
    w.f = enc.node_embed.in_features
    w.node_w, w.node_b = p(enc.node_embed.weight), p(enc.node_embed.bias)
    dep = getattr(enc, "depot_embed", None)
    w.depot_w = p(dep.weight) if dep is not None else None
    w.depot_b = p(dep.bias) if dep is not None else None
    assert len(enc.attention_layers) == vrpx.LAYERS, "libvrpx is built for 3 encoder layers"
    for i, layer in enumerate(enc.attention_layers):
        L = w.layer[i]
        a = layer.attention_layer
        L.in_proj_w, L.in_proj_b = p(a.in_proj_weight), p(a.in_proj_bias)
        L.out_proj_w, L.out_proj_b = p(a.out_proj.weight), p(a.out_proj.bias)
        L.bn1_w, L.bn1_b = p(layer.bn1.norm.weight), p(layer.bn1.norm.bias)
        L.bn1_mean, L.bn1_var = p(layer.bn1.norm.running_mean), p(layer.bn1.norm.running_var)
        L.ff0_w, L.ff0_b = p(layer.ff[0].weight), p(layer.ff[0].bias)
        L.ff2_w, L.ff2_b = p(layer.ff[2].weight), p(layer.ff[2].bias)
        L.bn2_w, L.bn2_b = p(layer.bn2.norm.weight), p(layer.bn2.norm.bias)
        L.bn2_mean, L.bn2_var = p(layer.bn2.norm.running_mean), p(layer.bn2.norm.running_var)
    return w
