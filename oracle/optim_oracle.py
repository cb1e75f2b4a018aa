"""TEST INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's densification bookkeeping
(/root/reference/scene/gaussian_model.py): add_densification_stats :1270-1276, densify_and_clone :1188-1222,
densify_and_split :1136-1186, prune_points / _prune_optimizer :1019-1057, cat_tensors_to_optimizer /
densification_postfix :1059-1134, densify_and_prune :1224-1250. PINNED: tests/golden/ref_model.npz holds the inputs and
outputs (all 13 parameter groups, both Adam moments, the statistics) of the reference's own densify_and_prune run on CPU
with its device="cuda" factories redirected and its torch.normal draws recorded (tests/golden/make_golden_model.py);
tests/test_model_golden_cpu.py requires this restatement to reproduce them bit for bit.
The Adam half needs no restatement: the reference's optimiser IS torch.optim.Adam (:769), which the tests run directly.
torch.normal(mean=0, std=stds) is taken as stds * z with caller-supplied z ~ N(0,1). Only tests/ may import this."""
import torch


def build_rotation(r):
    """utils/general_utils.py:117-149."""
    q = r / torch.sqrt((r * r).sum(-1, keepdim=True))
    R = torch.zeros((q.shape[0], 3, 3), dtype=q.dtype)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def add_densification_stats(stats, viewspace_grad, update_filter, weights):
    stats["weights_accum"] += weights
    stats["xyz_gradient_accum"][update_filter] += torch.norm(viewspace_grad[update_filter, :2], dim=-1, keepdim=True)
    stats["denom"][update_filter] += 1


def densify_and_prune(t, moments, stats, max_grad, min_opacity, extent, max_screen_size, max_grad_normal, percent_dense,
                      weights_threshold, z):
    """t: dict name -> [P,...] tensors ("xyz", "scaling", "rotation", "opacity" + anything else); moments: dict name ->
    (exp_avg, exp_avg_sq). Returns (new t, new moments, new stats)."""
    t = {k: v.clone() for k, v in t.items()}
    moments = {k: (a.clone(), b.clone()) for k, (a, b) in moments.items()}
    get_scaling = lambda: torch.exp(t["scaling"])
    grads = stats["xyz_gradient_accum"] / stats["denom"]
    grads_normal = stats["normal_gradient_accum"] / stats["denom"]
    grads[grads.isnan()] = 0.0
    grads_normal[grads_normal.isnan()] = 0.0
    weights_accum = stats["weights_accum"].clone()

    def cat(new):                       # cat_tensors_to_optimizer + densification_postfix
        nonlocal weights_accum
        n = new["xyz"].shape[0]
        for k in t:
            t[k] = torch.cat([t[k], new[k]], 0)
            if k in moments:
                a, b = moments[k]
                moments[k] = (torch.cat([a, torch.zeros_like(new[k])], 0), torch.cat([b, torch.zeros_like(new[k])], 0))
        weights_accum = torch.cat([weights_accum, torch.ones((n, 1))], 0)

    def prune(mask):                    # prune_points
        nonlocal weights_accum
        valid = ~mask
        for k in t:
            t[k] = t[k][valid]
            if k in moments:
                a, b = moments[k]
                moments[k] = (a[valid], b[valid])
        weights_accum = weights_accum[valid]

    # densify_and_clone
    sel = torch.norm(grads, dim=-1) >= max_grad
    sel = torch.logical_or(sel, torch.norm(grads_normal, dim=-1) >= max_grad_normal)
    sel = torch.logical_and(sel, torch.max(get_scaling(), dim=1).values <= percent_dense * extent)
    cat({k: v[sel] for k, v in t.items()})
    # densify_and_split (N = 2)
    n_init = t["xyz"].shape[0]
    padded = torch.zeros(n_init); padded[:grads.shape[0]] = grads.squeeze(-1)
    padded_n = torch.zeros(n_init); padded_n[:grads_normal.shape[0]] = grads_normal.squeeze(-1)
    sel = torch.logical_or(padded >= max_grad, padded_n >= max_grad_normal)
    sel = torch.logical_and(sel, torch.max(get_scaling(), dim=1).values > percent_dense * extent)
    stds = get_scaling()[sel].repeat(2, 1)
    samples = stds * z                  # torch.normal(mean=0, std=stds)
    rots = build_rotation(t["rotation"][sel]).repeat(2, 1, 1)
    new = {k: v[sel].repeat(*([2] + [1] * (v.dim() - 1))) for k, v in t.items()}
    new["xyz"] = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + t["xyz"][sel].repeat(2, 1)
    new["scaling"] = torch.log(get_scaling()[sel].repeat(2, 1) / (0.8 * 2))
    new["scaling"][:, -1] = -1e10
    n_sel = int(sel.sum())
    cat(new)
    prune(torch.cat((sel, torch.zeros(2 * n_sel, dtype=torch.bool))))
    # final prune
    mask = (torch.sigmoid(t["opacity"]) < min_opacity).squeeze(-1)
    mask = torch.logical_or(weights_accum[:, 0] < weights_threshold, mask)
    if max_screen_size:
        max_radii2D = torch.zeros(t["xyz"].shape[0])          # densification_postfix reset it (:1123)
        big_vs = max_radii2D > max_screen_size
        big_ws = get_scaling().max(dim=1).values > 0.1 * extent
        mask = torch.logical_or(torch.logical_or(mask, big_vs), big_ws)
    prune(mask)
    P2 = t["xyz"].shape[0]
    new_stats = {k: torch.zeros((P2, 1)) for k in ("weights_accum", "xyz_gradient_accum", "normal_gradient_accum", "denom")}
    return t, moments, new_stats
