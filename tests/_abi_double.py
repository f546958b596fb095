"""TEST DOUBLE of the C-ABI contract (include/latte_b200.h) in plain torch on the CPU.

Only for the host-logic tests that run without a GPU (world_size-2 gloo): it lets the
collective choreography and coefficient rules of latteclip_b200.loss / .prototypes be
checked against the golden gloo run of the reference.  It is never used by the product
path (latteclip_b200 has no CPU fallback) and is not the oracle either: it restates the
header's formulas, not the reference's code."""

import torch


def rank_sweep_supported(dtype, dim):
    return True      # exercise the one-sweep multi-rank host logic on the CPU


def clip_fwd_rows(img_loc, txt_all, label_offset, logit_scale):
    s = logit_scale.detach().double().reshape(())
    il, ta = img_loc.detach().double(), txt_all.detach().double()
    n = il.shape[0]
    sm = s * il @ ta.T                                   # [n, N]
    row_lse = torch.logsumexp(sm, 1)
    label_logit = sm[torch.arange(n), torch.arange(n) + label_offset]
    row_nll = row_lse - label_logit
    x2 = sm * LOG2E
    m = x2.max(0).values
    col_ml = torch.stack([m, torch.exp2(x2 - m[None, :]).sum(0)], dim=1)     # [N, 2] base 2
    return torch.cat([col_ml.reshape(-1), row_lse, row_nll, label_logit]).float()


def clip_fwd_cols(gathered, img_all, txt_all, n_loc, label_offset, logit_scale):
    world, big_n = gathered.shape[0], img_all.shape[0]
    gd = gathered.double()
    ml = gd[:, :2 * big_n].reshape(world, big_n, 2)
    row_lse_all = gd[:, 2 * big_n:2 * big_n + n_loc].reshape(-1)
    row_nll_all = gd[:, 2 * big_n + n_loc:2 * big_n + 2 * n_loc].reshape(-1)
    label_logit_all = gd[:, 2 * big_n + 2 * n_loc:].reshape(-1)
    m = ml[:, :, 0].max(0).values
    big_l = (ml[:, :, 1] * torch.exp2(ml[:, :, 0] - m[None, :])).sum(0)
    col_lse_all = (m + torch.log2(big_l)) / LOG2E
    col_nll_all = col_lse_all - label_logit_all
    own = slice(label_offset, label_offset + n_loc)
    loss = (row_nll_all[own].mean() + col_nll_all[own].mean()) / 2
    return (row_lse_all.float(), row_nll_all.float(), col_lse_all.float(), col_nll_all.float(),
            loss.float().reshape(1), _stats(row_lse_all, col_lse_all, row_nll_all, col_nll_all))


def _stats(row_lse, col_lse, row_nll, col_nll):
    both = torch.cat([row_lse, col_lse])
    return torch.stack([both.min(), both.max(), torch.cat([row_nll, col_nll]).max(),
                        both.new_zeros(())]).float()


LOG2E = 1.4426950408889634


def clip_fwd(img_loc, txt_loc, img_all, txt_all, label_offset, logit_scale, with_nll=False,
             with_stats=False):
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
    out = (row_lse.float(), col_lse.float(), loss.float().reshape(1))
    if with_nll:
        out = out + ((row_lse - diag_r).float(), (col_lse - diag_c).float())
    if with_stats:
        out = out + (_stats(row_lse, col_lse, row_lse - diag_r, col_lse - diag_c),)
    return out


def clip_bwd(img_loc, txt_loc, img_all, txt_all, label_offset, logit_scale, row_lse_all,
             col_lse_all, grad_loss, grad_mult, cross_terms, grad_dtype=None, row_nll_all=None,
             col_nll_all=None, partial=False, comm=None, phases=3, lse_stats=None, out=None):
    s = logit_scale.detach().double().reshape(())
    if img_all is None:
        img_all = img_loc          # one-sweep modes read this rank's images only
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
    if partial:
        # one-sweep mode: G of this rank's rows x all columns; the text gradient is the partial
        # G^T @ img_loc over every column, d_scale covers rows of this rank x all columns
        sm = s * il @ ta.T
        g = torch.exp(sm - row_all[idx][:, None]) + torch.exp(sm - col_all[None, :])
        g[torch.arange(n), idx] -= 2.0
        d_img = coef * s * g @ ta
        d_part = coef * s * g.T @ il
        ds = coef * (g * (sm / s)).sum()
        return d_img.to(img_loc.dtype), d_part.float(), ds.float().reshape(1)
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


# ------------------------------------------------------------------ SigLIP entry points
def siglip_supported(dtype, dim):
    return True


def _siglip_terms(img_loc, txt_all, label_offset, logit_scale, logit_bias):
    s = logit_scale.detach().double().reshape(())
    b = logit_bias.detach().double().reshape(()) if logit_bias is not None else torch.zeros((), dtype=torch.float64)
    il, ta = img_loc.detach().double(), txt_all.detach().double()
    n = il.shape[0]
    dots = il @ ta.T
    z = s * dots + b
    lab = -torch.ones_like(z)
    lab[torch.arange(n), torch.arange(n) + label_offset] = 1.0
    return il, ta, dots, z, lab, s


def siglip_fwd(img_loc, txt_all, label_offset, logit_scale, logit_bias=None):
    il, ta, dots, z, lab, s = _siglip_terms(img_loc, txt_all, label_offset, logit_scale, logit_bias)
    return (torch.nn.functional.softplus(-lab * z).sum() / il.shape[0]).float().reshape(1)


def siglip_bwd(img_loc, txt_all, label_offset, logit_scale, logit_bias, grad_loss, grad_dtype=None,
               partial=False, peer_ptrs=None):
    assert peer_ptrs is None
    il, ta, dots, z, lab, s = _siglip_terms(img_loc, txt_all, label_offset, logit_scale, logit_bias)
    n = il.shape[0]
    g = (torch.sigmoid(z) - (lab > 0).double()) * (grad_loss.detach().double().reshape(()) / n)
    gdt = img_loc.dtype if grad_dtype is None else grad_dtype
    d_img = (s * g @ ta).to(gdt)
    d_txt = s * g.T @ il
    d_txt = d_txt.float() if partial else d_txt.to(gdt)
    return d_img, d_txt, (g * dots).sum().float().reshape(1), g.sum().float().reshape(1)


# ------------------------------------------------------------------ prototype entry points
# (host-logic tests of latteclip_b200.train_step: restated from include/latte_b200.h)
def nxc_multi_supported(x, num_classes):
    return False     # the per-product entry points below carry the host logic


def normalize_rows(x):
    return torch.nn.functional.normalize(x.detach().float(), dim=1)


def nxc_argmax_margin(x, protos, scale=1.0, row_index=None, want_argmax=True, want_margin=True,
                      want_top1=False):
    xs = x.detach().double()
    if row_index is not None:
        xs = xs[row_index]
    logits = scale * xs @ protos.detach().double().T
    top2 = logits.topk(2, dim=1).values
    am = logits.argmax(1) if want_argmax else None
    mg = (top2[:, 0] - top2[:, 1]).float() if want_margin else None
    return am, mg, (top2[:, 0].float() if want_top1 else None)


def _mix_coeffs(w_lbl, w_lbl_zs, w_img, w_grp, alpha, axis, b, d):
    tot, tot_z = (w_lbl + w_img + w_grp).double(), (w_lbl_zs + w_img + w_grp).double()
    wl = w_lbl.double()[None, :].expand(b, d) if axis == "quirk" else w_lbl.double()[:, None].expand(b, d)
    return tot, tot_z, wl


def mix_ema_fwd(class_text, per_image, per_group, bank, preds, zs, w_lbl, w_lbl_zs, w_img, w_grp,
                alpha, label_axis):
    b, d = per_image.shape
    tot, tot_z, wl = _mix_coeffs(w_lbl, w_lbl_zs, w_img, w_grp, alpha, label_axis, b, d)
    ct, pi, pg, bk = (x.detach().double() for x in (class_text, per_image, per_group, bank))
    common = pi * w_img.double()[:, None] + pg * w_grp.double()[:, None]
    mix_ft = (wl * ct[preds] + common) / tot[:, None]
    mix_zs = (wl * ct[zs] + common) / tot_z[:, None]
    t_ft = bk[preds] + alpha * (mix_ft - bk[preds])
    t_zs = bk[zs] + alpha * (mix_zs - bk[zs])
    return t_ft.to(per_image.dtype), t_zs.to(per_image.dtype)


def mix_ema_bwd(d_t_ft, d_t_zs, preds, zs, w_lbl, w_lbl_zs, w_img, w_grp, alpha, label_axis,
                num_classes, want_bank=False):
    b, d = d_t_ft.shape
    tot, tot_z, wl = _mix_coeffs(w_lbl, w_lbl_zs, w_img, w_grp, alpha, label_axis, b, d)
    gf, gz = d_t_ft.detach().double(), d_t_zs.detach().double()
    af, az = alpha * gf / tot[:, None], alpha * gz / tot_z[:, None]
    d_ct = torch.zeros(num_classes, d, dtype=torch.float64)
    d_ct.index_add_(0, preds, wl * af)
    d_ct.index_add_(0, zs, wl * az)
    d_pi = (af + az) * w_img.double()[:, None]
    d_pg = (af + az) * w_grp.double()[:, None]
    d_bank = None
    if want_bank:
        d_bank = torch.zeros(num_classes, d, dtype=torch.float64)
        d_bank.index_add_(0, preds, (1 - alpha) * gf)
        d_bank.index_add_(0, zs, (1 - alpha) * gz)
        d_bank = d_bank.float()
    return d_ct.float(), d_pi.to(d_t_ft.dtype), d_pg.to(d_t_ft.dtype), d_bank


def install(_lib):
    """Point every entry point the host logic uses at this double."""
    for name in ("clip_fwd", "clip_bwd", "clip_fwd_rows", "clip_fwd_cols", "bank_accumulate",
                 "bank_finalize", "nxc_multi_supported", "normalize_rows", "nxc_argmax_margin",
                 "mix_ema_fwd", "mix_ema_bwd"):
        setattr(_lib, name, globals()[name])
