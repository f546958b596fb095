"""TEST DOUBLE of the C-ABI contract (include/latte_b200.h) in plain torch on the CPU.

Only for the host-logic tests that run without a GPU (world_size-2 gloo): it lets the
collective choreography and coefficient rules of latteclip_b200.loss / .prototypes be
checked against the golden gloo run of the reference.  It is never used by the product
path (latteclip_b200 has no CPU fallback) and is not the oracle either: it restates the
header's formulas, not the reference's code."""

import torch


def clip_fwd(img_loc, txt_loc, img_all, txt_all, label_offset, logit_scale):
    s = logit_scale.detach().double().reshape(())
    il, tl, ia, ta = (x.detach().double() for x in (img_loc, txt_loc, img_all, txt_all))
    n = il.shape[0]
    s_r = s * il @ ta.T
    s_c = s * tl @ ia.T
    row_lse = torch.logsumexp(s_r, 1)
    col_lse = torch.logsumexp(s_c, 1)
    idx = torch.arange(n) + label_offset
    diag_r = s_r[torch.arange(n), idx]
    diag_c = s_c[torch.arange(n), idx]
    loss = ((row_lse - diag_r).mean() + (col_lse - diag_c).mean()) / 2
    return row_lse.float(), col_lse.float(), loss.float().reshape(1)


def clip_bwd(img_loc, txt_loc, img_all, txt_all, label_offset, logit_scale, row_lse_all,
             col_lse_all, grad_loss, grad_mult, cross_terms):
    s = logit_scale.detach().double().reshape(())
    il, tl, ia, ta = (x.detach().double() for x in (img_loc, txt_loc, img_all, txt_all))
    n = il.shape[0]
    idx = torch.arange(n) + label_offset
    coef = grad_loss.detach().double().reshape(()) * grad_mult / (2 * n)
    ca, cb, cd = (1.0, 1.0, 2.0) if cross_terms else (1.0, 0.0, 1.0)
    row_all, col_all = row_lse_all.double(), col_lse_all.double()

    def side(x_loc, y_all, lse_a_all, lse_b_all):
        sm = s * x_loc @ y_all.T
        ea = torch.exp(sm - lse_a_all[idx][:, None])
        eb = torch.exp(sm - lse_b_all[None, :])
        g = ca * ea + cb * eb
        g[torch.arange(n), idx] -= cd
        dx = coef * s * g @ y_all
        ea[torch.arange(n), idx] -= 1.0
        ds = coef * (ea * (sm / s)).sum()
        return dx, ds
    d_img, ds_a = side(il, ta, row_all, col_all)
    d_txt, ds_b = side(tl, ia, col_all, row_all)
    return d_img.to(img_loc.dtype), d_txt.to(img_loc.dtype), (ds_a + ds_b).float().reshape(1)


def bank_accumulate(t_ft, t_zs, preds, zs, num_classes):
    d = t_ft.shape[1]
    sums = torch.zeros(num_classes, d)
    sums.index_add_(0, zs, t_zs.float())
    sums.index_add_(0, preds, t_ft.float())
    counts = (torch.bincount(zs, minlength=num_classes) + torch.bincount(preds, minlength=num_classes)).float()
    return sums, counts


def bank_finalize(sums, counts, bank):
    m = counts > 0
    mean = sums[m] / counts[m][:, None]
    bank[m] = torch.nn.functional.normalize(mean, dim=1)
    return bank
