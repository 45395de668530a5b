"""optim.sgd_list (src/optim/optim_sgd.lua:23-99), default branch, over the parameter proxies.

The reference evaluates `opfunc`, then per group: prints the norms (:49), clips the gradient to L2 norm 5
(:50-52) and applies `y:add(-clr, dfdy)` (:90).  Weight decay / momentum / per-parameter learning rates are
accepted only at their defaults (0 / nil), as the reference's caller never sets them (model.lua:700).
"""


def sgd_list(opfunc, x, config, state=None, verbose=False):
    state = config if state is None else state
    lr = config.get("learningRate", 1e-3)
    lrd = config.get("learningRateDecay", 0)
    assert config.get("weightDecay", 0) == 0 and config.get("momentum", 0) == 0, "only the reference defaults"
    fx, dfdx, stats = opfunc(x)
    for i in range(len(x)):
        st = state.setdefault(i + 1, {})
        nevals = st.get("evalCounter", 0)
        y, dfdy = x[i], dfdx[i]
        gnorm = dfdy.norm()
        if verbose:
            print("i: %d, param norm: %f, grad norm: %f" % (i + 1, y.norm(), gnorm))
        if gnorm > 5:
            dfdy.mul(5.0 / gnorm)
        clr = lr / (1 + nevals * lrd)
        y.add(-clr, dfdy)
        st["evalCounter"] = nevals + 1
    return x, [fx], stats
