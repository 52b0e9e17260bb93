"""A torch restatement of the ALGEBRA the CUDA rollout kernel uses (csrc/rollout.cu phases P1-P4 on the packed
arrays of vrpx/packing.py).  Test helper only: it lets the CPU suite prove that the folded weights reproduce the
reference decoder before any GPU time is spent, and documents the kernel's math in ten lines."""
import torch

H, E = 8, 128


def kernel_logits(packed, h, mask, t, first_idx, last_idx, load=None):
    """packed: dict of f32 tensors from fold_decoder; h (B,N,E); mask (B,N) f32 0/1; t step index;
    first_idx/last_idx (B,) long (ignored at t == 0).  Returns masked pointer logits (B,N)."""
    B, N, _ = h.shape
    g = h.sum(1) * (1.0 / N)
    qg = g @ packed["ag_t"] + packed["a_c"]                       # prologue (GEMM-A, EPI 0)
    if t == 0:
        qt = qg + packed["a_q0"]
    else:
        ar = torch.arange(B)
        if packed["af_t"] is not None:
            qg = qg + h[ar, first_idx] @ packed["af_t"]           # step-1 fold of `first` (EPI 1)
        qt = qg + h[ar, last_idx] @ packed["al_t"]                # GEMM-A, EPI 2
    if load is not None:
        qt = qt + load[:, None] * packed["a_load"]
    qt = qt.view(B, H, E)
    s = torch.einsum("bhe,bne->bhn", qt, h)
    rows = (torch.arange(B)[:, None] * H + torch.arange(H)[None, :]) % B
    s = s + mask[rows]
    p = torch.softmax(s, -1)
    c = torch.einsum("bhn,bne->bhe", p, h).reshape(B, H * E)
    qh = c @ packed["m_t"] + packed["m_c"]                        # GEMM-B
    u = 10.0 * torch.tanh(torch.einsum("be,bne->bn", qh, h))
    return u.masked_fill(mask.bool(), float("-inf"))
