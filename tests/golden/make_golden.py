"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules (`src.flows`, `src.models`) from
/root/reference (with an in-memory `torchdiffeq` stub, because
`src/flows/__init__.py:9` eagerly imports the CNF path which is out of scope),
builds each hot-path layer/model with a fixed seed, perturbs the weights so no
layer is the identity (the reference zero-initialises final layers,
coupling_layer.py:108-111, spline_coupling_layer.py:319-323), runs
forward/inverse on fixed inputs on CPU in fp32, and stores
{ctor, state_dict, inputs, outputs} as small `.pt` files next to this script.

The reference has no golden vectors of its own; these files are the
parity pin for `oracle/flows_oracle.py` and, through it, for the CUDA kernels.
Nothing reads /root/reference at test time.
"""
import os
import sys
import types

import torch

REF = os.environ.get("NF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    stub = types.ModuleType("torchdiffeq")
    stub.odeint = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("torchdiffeq stub"))
    stub.odeint_adjoint = stub.odeint
    sys.modules["torchdiffeq"] = stub
    # make sure the repo's own drop-in `src` shim cannot shadow the reference
    sys.path[:] = [p for p in sys.path if "normalizing-flows-study_b200" not in p]
    sys.path.insert(0, REF)
    import src.flows as F  # noqa
    import src.models as M  # noqa
    return F, M


def perturb(module, g, sigma):
    """reference init + sigma*randn on every float parameter; non-trivial BN stats."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            p.add_(sigma * torch.randn(p.shape, generator=g))
        for n, b in module.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(0.2 * torch.randn(b.shape, generator=g))
            elif n.endswith("running_var"):
                b.copy_(0.5 + torch.rand(b.shape, generator=g))


def clone_sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def run_both(mod, x):
    with torch.no_grad():
        f, fl = mod.forward(x)
        i, il = mod.inverse(x)
    return {"fwd": f.clone(), "fwd_ld": torch.as_tensor(fl).clone(),
            "inv": i.clone(), "inv_ld": torch.as_tensor(il).clone()}


def stress_rows(D, g):
    """inputs hitting clamps, tails and the NaN/Inf scrubs."""
    x = torch.randn(24, D, generator=g) * 2.0
    x[0] = 0.0
    x[1] = 1e-6
    x[2] = 1e3
    x[3] = -1e3
    x[4] = 1e10
    x[5, 0] = float("nan")
    x[6, -1] = float("inf")
    x[7, 0] = float("-inf")
    x[8] = 5.0          # exactly +bound for the spline
    x[9] = -5.0         # exactly -bound
    x[10] = 5.0000005
    x[11] = 7.5         # outside
    return x


def main():
    RF, RM = _import_reference()
    g = torch.Generator().manual_seed(1234)
    cases = {}

    def mask(D, kind):
        m = torch.zeros(D)
        if kind == "alt":
            m[::2] = 1
        else:
            m[: D // 2] = 1
        return m

    # ---- a1/a2 affine coupling: eval + train (batch stats) ---------------------
    for D, H, mk in ((4, 16, "alt"), (4, 16, "half"), (5, 8, "alt"), (2, 64, "half")):
        torch.manual_seed(7)
        layer = RF.CouplingLayer(D, H, mask(D, mk))
        perturb(layer, g, 0.3)
        x = torch.cat([torch.randn(40, D, generator=g), stress_rows(D, g)])
        layer.eval()
        sd = clone_sd(layer)
        out = run_both(layer, x)
        cases[f"coupling_eval_D{D}_H{H}_{mk}"] = dict(kind="coupling", D=D, H=H, sd=sd, x=x, **out)
        # train mode: batch statistics + running-stat side effects (finite rows only)
        xt = torch.randn(48, D, generator=g)
        layer.train()
        sd0 = clone_sd(layer)
        with torch.no_grad():
            f, fl = layer.forward(xt)
        sd1 = clone_sd(layer)
        cases[f"coupling_train_D{D}_H{H}_{mk}"] = dict(kind="coupling_train", D=D, H=H, sd=sd0, sd_after=sd1,
                                                      x=xt, fwd=f.clone(), fwd_ld=fl.clone())

    # ---- a4-a6 spline coupling ------------------------------------------------
    for D, H, K, mk, extra in ((4, 16, 10, "alt", {}), (4, 16, 8, "half", {}), (3, 8, 4, "alt", {}),
                               (5, 16, 10, "half", {}), (2, 64, 8, "half", {}),
                               (4, 16, 6, "alt", dict(bound=3.0)),
                               (4, 16, 10, "half", dict(data_min=-2.0, data_max=6.0))):
        torch.manual_seed(11)
        layer = RF.SplineCouplingLayer(D, H, mask(D, mk), num_bins=K, **extra)
        perturb(layer, g, 0.4)
        layer.eval()
        x = torch.cat([torch.randn(40, D, generator=g) * 2.5, stress_rows(D, g)])
        tag = f"spline_D{D}_H{H}_K{K}_{mk}" + ("_" + "_".join(extra) if extra else "")
        cases[tag] = dict(kind="spline", D=D, H=H, K=K, extra=extra, sd=clone_sd(layer), x=x, **run_both(layer, x))
        # all-outside batch: early-return branch (spline_coupling_layer.py:200-201)
        xo = 6.0 + torch.rand(8, D, generator=g)
        cases[tag + "_outside"] = dict(kind="spline", D=D, H=H, K=K, extra=extra, sd=clone_sd(layer), x=xo,
                                       **run_both(layer, xo))
        # bare bounded-spline function on exact knot hits
        if not extra:
            Dt = int((layer.mask == 0).sum())
            uw = torch.randn(16, Dt, K, generator=g) * 2
            uh = torch.randn(16, Dt, K, generator=g) * 2
            ud = torch.randn(16, Dt, K - 1, generator=g) * 2
            xin = torch.rand(16, Dt, generator=g) * 12 - 6
            xin[0] = -5.0
            xin[1] = 5.0
            with torch.no_grad():
                of, lf = layer._rational_quadratic_spline(xin, uw, uh, ud, inverse=False)
                oi, li = layer._rational_quadratic_spline(xin, uw, uh, ud, inverse=True)
            cases[tag + "_rqs"] = dict(kind="rqs_bounded", K=K, bound=5.0, x=xin, uw=uw, uh=uh, ud=ud,
                                       fwd=of, fwd_ld=lf, inv=oi, inv_ld=li)

    # ---- a7 public spline on [0,1] -----------------------------------------
    for K in (4, 8, 10):
        n = 96
        w = torch.randn(n, K, generator=g) * 1.5
        h = torch.randn(n, K, generator=g) * 1.5
        d = torch.randn(n, K - 1, generator=g) * 1.5
        x = torch.rand(n, generator=g)
        x[0], x[1], x[2], x[3], x[4] = 0.0, 1.0, -0.5, 1.5, 0.5
        of, lf = RF.rational_quadratic_spline(x, w, h, d, inverse=False)
        oi, li = RF.rational_quadratic_spline(x, w, h, d, inverse=True)
        cases[f"rqs_unit_K{K}"] = dict(kind="rqs_unit", K=K, x=x, w=w, h=h, d=d, fwd=of, fwd_ld=lf, inv=oi, inv_ld=li)
    # (params with >2 dims raise inside the reference: rational_quadratic_spline.py:46-54 flattens
    #  the knots but not the inputs, so [N] inputs with [N,K] params is the only supported shape)

    # ---- a8-a13 MADE / MAF / IAF ----------------------------------------
    for D, H in ((1, 8), (2, 16), (3, 8), (5, 32), (10, 24), (16, 64)):
        for name, cls in (("maf", RF.MaskedAutoregressiveFlow), ("iaf", RF.InverseAutoregressiveFlow)):
            torch.manual_seed(3)
            layer = cls(D, H)
            perturb(layer, g, 0.25)
            layer.eval()
            x = torch.cat([torch.randn(24, D, generator=g), stress_rows(D, g)[:8] * 0.01 * 100])
            sd = clone_sd(layer)
            cases[f"{name}_D{D}_H{H}"] = dict(kind=name, D=D, H=H, sd=sd, x=x, **run_both(layer, x),
                                              degrees=torch.as_tensor(layer.conditioner.m[0]))

    # ---- a14-a16 stacks ------------------------------------------------------
    torch.manual_seed(5)
    m = RM.RealNVP(2, 8, 64)
    perturb(m, g, 0.05)
    m.eval()
    x = torch.randn(256, 2, generator=g)
    cases["realnvp_2_8_64"] = dict(kind="realnvp", D=2, L=8, H=64, bn=False, sd=clone_sd(m), x=x, **run_both(m, x))

    torch.manual_seed(5)
    m = RM.RealNVP(4, 4, 16, batch_norm_between_layers=True)
    perturb(m, g, 0.1)
    m.eval()
    x = torch.randn(64, 4, generator=g)
    cases["realnvp_4_4_16_bn"] = dict(kind="realnvp", D=4, L=4, H=16, bn=True, sd=clone_sd(m), x=x, **run_both(m, x))
    # train mode forward: between-layer BN running-stat update (normalizing_flow_model.py:74-79)
    m.train()
    sd0 = clone_sd(m)
    with torch.no_grad():
        f, fl = m.forward(x)
    cases["realnvp_4_4_16_bn_train"] = dict(kind="realnvp_train", D=4, L=4, H=16, bn=True, sd=sd0,
                                            sd_after=clone_sd(m), x=x, fwd=f.clone(), fwd_ld=fl.clone())

    torch.manual_seed(5)
    m = RM.RealNVPSpline(2, 8, 64)
    perturb(m, g, 0.05)
    m.eval()
    x = torch.randn(256, 2, generator=g) * 1.5
    cases["realnvpspline_2_8_64"] = dict(kind="realnvpspline", D=2, L=8, H=64, K=10, bn=False, sd=clone_sd(m), x=x,
                                         **run_both(m, x))

    torch.manual_seed(5)
    m = RM.RealNVPSpline(6, 4, 32, batch_norm_between_layers=True)
    perturb(m, g, 0.1)
    m.eval()
    x = torch.randn(64, 6, generator=g) * 1.5
    cases["realnvpspline_6_4_32_bn"] = dict(kind="realnvpspline", D=6, L=4, H=32, K=10, bn=True, sd=clone_sd(m), x=x,
                                            **run_both(m, x))

    # config-2 shape: 8 spline layers, K=8, D=2, H=64 built by hand
    torch.manual_seed(5)
    masks = [mask(2, "half") if i % 2 == 0 else 1 - mask(2, "half") for i in range(8)]
    m = RM.NormalizingFlowModel([RF.SplineCouplingLayer(2, 64, mk, num_bins=8) for mk in masks])
    perturb(m, g, 0.05)
    m.eval()
    x = torch.randn(256, 2, generator=g) * 1.5
    cases["splinestack_2_8_64_K8"] = dict(kind="splinestack", D=2, L=8, H=64, K=8, bn=False, sd=clone_sd(m), x=x,
                                          **run_both(m, x))

    # mixed stack + SequentialFlow
    torch.manual_seed(5)
    m = RM.NormalizingFlowModel([RF.CouplingLayer(4, 16, mask(4, "alt")), RF.MaskedAutoregressiveFlow(4, 16),
                                 RF.SplineCouplingLayer(4, 16, mask(4, "half")), RF.InverseAutoregressiveFlow(4, 16)],
                                batch_norm_between_layers=True)
    perturb(m, g, 0.15)
    m.eval()
    x = torch.randn(48, 4, generator=g)
    cases["mixed_4"] = dict(kind="mixed", D=4, H=16, bn=True, sd=clone_sd(m), x=x, **run_both(m, x),
                            specs=[dict(kind="coupling"), dict(kind="maf"), dict(kind="spline", num_bins=10),
                                   dict(kind="iaf")])
    torch.manual_seed(5)
    m = RF.SequentialFlow([RF.MaskedAutoregressiveFlow(3, 16), RF.InverseAutoregressiveFlow(3, 16)])
    perturb(m, g, 0.15)
    m.eval()
    x = torch.randn(32, 3, generator=g)
    cases["sequential_3"] = dict(kind="sequential", D=3, H=16, sd=clone_sd(m), x=x, **run_both(m, x),
                                 specs=[dict(kind="maf"), dict(kind="iaf")])

    total = 0
    for name, blob in cases.items():
        path = os.path.join(HERE, name + ".pt")
        torch.save(blob, path)
        total += os.path.getsize(path)
    print(f"wrote {len(cases)} golden files, {total/1e6:.2f} MB, torch {torch.__version__}")


if __name__ == "__main__":
    main()
