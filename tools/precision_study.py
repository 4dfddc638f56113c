"""Which operand roundings of the attention block cost how much against the fp32 reference?

CPU study (test infrastructure, uses the oracle's layout helpers): every tensor-core operand of the block can be
rounded to one of: 'f32' (exact), 'f16' (one fp16 value), 'bf16x2' (bf16 hi + lo = 16 significant bits),
'f16x2' (fp16 hi + lo = 22 bits).  Prints max |err|, rms and the number of elements outside 1e-3 rel / 1e-4 abs.
"""
import itertools, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_ops as R

def rounder(kind):
    if kind == "f32":
        return lambda t: t
    if kind == "f16":
        return lambda t: t.half().float()
    if kind == "bf16":
        return lambda t: t.bfloat16().float()
    if kind == "bf16x2":
        def f(t):
            hi = t.bfloat16().float()
            return hi + (t - hi).bfloat16().float()
        return f
    if kind == "f16x2":
        def f(t):
            hi = t.half().float()
            return hi + (t - hi).half().float()
        return f
    raise ValueError(kind)

STAGES = ["x", "wqkv", "q", "k", "v", "p", "o", "wproj"]

def block(xw, p, heads, ws, kinds):
    r = {s: rounder(kinds.get(s, "f32")) for s in STAGES}
    K, N, C = xw.shape
    d = C // heads
    scale = d ** -0.5
    w = p["qkv_w"].clone(); w[:C] *= scale
    b = p["qkv_b"].clone(); b[:C] *= scale
    qkv = (r["x"](xw) @ r["wqkv"](w).t() + b).reshape(K, N, 3, heads, d)
    q = r["q"](qkv[:, :, 0].permute(0, 2, 1, 3))
    k = r["k"](qkv[:, :, 1].permute(0, 2, 1, 3))
    v = r["v"](qkv[:, :, 2].permute(0, 2, 1, 3))
    s_ = torch.einsum("khid,khjd->khij", q, k) + R.expand_bias(p["table"], ws)[None]
    e = torch.exp(s_ - s_.amax(-1, keepdim=True))
    o = torch.einsum("khij,khjd->kihd", r["p"](e), v) / e.sum(-1).permute(0, 2, 1)[..., None]
    o = r["o"](o.reshape(K, N, C))
    return o @ r["wproj"](p["proj_w"]).t() + p["proj_b"]

def main():
    torch.manual_seed(0)
    C, heads, ws = 192, 8, 8
    K = 768
    lin1 = torch.nn.Linear(C, 3 * C); lin2 = torch.nn.Linear(C, C)
    table = torch.nn.init.trunc_normal_(torch.zeros((2 * ws - 1) ** 2, heads), std=.02)
    p = dict(qkv_w=lin1.weight.detach(), qkv_b=lin1.bias.detach(), proj_w=lin2.weight.detach(),
             proj_b=lin2.bias.detach(), table=table)
    xw = torch.randn(K, ws * ws, C)
    ref64 = block(xw.double(), {k: v.double() for k, v in p.items()}, heads, ws, {}).float()
    ref = block(xw, p, heads, ws, {})
    out_ref = xw + ref
    def report(name, kinds):
        y = block(xw, p, heads, ws, kinds)
        err = (y - ref).abs()
        tol = 1e-4 + 1e-3 * out_ref.abs()
        print(f"{name:44s} max {err.max():.2e} rms {err.pow(2).mean().sqrt():.2e} "
              f"viol {(err > tol).sum().item():8d} worst err/tol {(err / tol).max():.2f}")
    print("fp32 vs fp64:", (ref - ref64).abs().max().item())
    report("all f16", {s: "f16" for s in STAGES})
    for s in STAGES:
        report(f"only {s} f16", {s: "f16"})
    for kind in ("bf16x2", "f16x2"):
        report(f"all {kind}", {s: kind for s in STAGES})
        for s in STAGES:
            kk = {t: kind for t in STAGES}; kk[s] = "f16"
            report(f"all {kind} but {s} f16", kk)
    # candidate mixes
    report("x,w*,o split(f16x2); q,k,v,p f16", dict(x="f16x2", wqkv="f16x2", o="f16x2", wproj="f16x2", q="f16", k="f16", v="f16", p="f16"))
    report("x,w*,o,v split; q,k,p f16", dict(x="f16x2", wqkv="f16x2", o="f16x2", wproj="f16x2", q="f16", k="f16", v="f16x2", p="f16"))
    report("x,w*,o,v,p split; q,k f16", dict(x="f16x2", wqkv="f16x2", o="f16x2", wproj="f16x2", q="f16", k="f16", v="f16x2", p="f16x2"))
    report("bf16x2: x,w*,o,v,p split; q,k f16", dict(x="bf16x2", wqkv="bf16x2", o="bf16x2", wproj="bf16x2", q="f16", k="f16", v="bf16x2", p="bf16x2"))

if __name__ == "__main__":
    main()


def policy_block(xw, p, heads, ws, qk="f16", rest="f16x2", pk=None, scale_in=1.0):
    """q/k columns: single pass in `qk`; v columns, P, O, Wproj in `rest` (pk overrides P)."""
    rq, rr = rounder(qk), rounder(rest)
    rp = rounder(pk) if pk else rr
    K, N, C = xw.shape
    d = C // heads
    scale = d ** -0.5
    w = p["qkv_w"].clone(); w[:C] *= scale
    b = p["qkv_b"].clone(); b[:C] *= scale
    qk_ = rq(xw) @ rq(w[:2 * C]).t() + b[:2 * C]
    v_ = rr(xw) @ rr(w[2 * C:]).t() + b[2 * C:]
    q = rq(qk_[..., :C].reshape(K, N, heads, d).permute(0, 2, 1, 3))
    k = rq(qk_[..., C:].reshape(K, N, heads, d).permute(0, 2, 1, 3))
    v = rr(v_.reshape(K, N, heads, d).permute(0, 2, 1, 3))
    s_ = torch.einsum("khid,khjd->khij", q, k) + R.expand_bias(p["table"], ws)[None]
    e = torch.exp(s_ - s_.amax(-1, keepdim=True))
    o = torch.einsum("khij,khjd->kihd", rp(e), v) / e.sum(-1).permute(0, 2, 1)[..., None]
    o = rr(o.reshape(K, N, C))
    return o @ rr(p["proj_w"]).t() + p["proj_b"]


def main2():
    C, heads, ws = 192, 8, 8
    for seed, sc in ((0, 1.0), (1, 1.0), (2, 4.0), (3, 0.25), (4, 16.0)):
        torch.manual_seed(seed)
        K = 768
        lin1 = torch.nn.Linear(C, 3 * C); lin2 = torch.nn.Linear(C, C)
        table = torch.nn.init.trunc_normal_(torch.zeros((2 * ws - 1) ** 2, heads), std=.02)
        p = dict(qkv_w=lin1.weight.detach(), qkv_b=lin1.bias.detach(), proj_w=lin2.weight.detach(),
                 proj_b=lin2.bias.detach(), table=table)
        xw = torch.randn(K, ws * ws, C) * sc
        ref = block(xw, p, heads, ws, {})
        out_ref = xw + ref
        tol = 1e-4 + 1e-3 * out_ref.abs()
        for name, kw in (("all f16", dict(qk="f16", rest="f16")),
                         ("qk f16, rest f16x2", dict(qk="f16", rest="f16x2")),
                         ("qk f16, rest f16x2, p f16", dict(qk="f16", rest="f16x2", pk="f16")),
                         ("qk f16, rest bf16x2", dict(qk="f16", rest="bf16x2")),
                         ("all f16x2", dict(qk="f16x2", rest="f16x2"))):
            err = (policy_block(xw, p, heads, ws, **kw) - ref).abs()
            print(f"seed {seed} x*{sc:<5} {name:28s} max {err.max():.2e} rms {err.pow(2).mean().sqrt():.2e} "
                  f"viol {(err > tol).sum().item():7d} worst err/tol {(err / tol).max():.2f}")

if __name__ == "__main__":
    main2()
