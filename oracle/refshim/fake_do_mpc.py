"""Probing stand-in for `do_mpc` (TEST INFRASTRUCTURE, fixture generation only).

do-mpc / CasADi / IPOPT are not installable here, so the reference's
position_control/mpc_cbf.py cannot SOLVE anything in this container.  What it can do is
STATE its problem: with casadi replaced by the numeric stand-in (fake_casadi.py), every
expression MPCCBF builds in create_model / create_mpc is evaluated on the spot.  This module
hands the reference numeric "variables" (the probe point set in PROBE) and records what the
reference's own code passes to do-mpc:

    set_rhs('x', .)            -> x_next at the probe (x, u)                      mpc_cbf.py:135-141
    set_expression('cost', .)  -> stage / terminal cost at the probe             :143-145, 175-178
    set_nl_cons('cbf_i', ., 0) -> minus the CBF constraint of obstacle slot i     :295-325
    set_rterm(u=R), bounds[...], set_param(n_horizon, t_step, n_robust, ...)      :164-232
    set_tvp_fun(f)             -> f(0) shows the goal padding / dummy obstacles / alphas   :261-293

What it cannot show is what do-mpc does with them (sum of lterm over the horizon + mterm,
rterm as a penalty on u_k - u_{k-1}, nl_cons at every stage but the terminal one): that
transcription stays as documented in SURVEY.md 8a and is marked unpinned.
"""
import types

import numpy as np

PROBE = {"_x": {}, "_u": {}, "_tvp": {}}      # var_type -> {var_name: ndarray}; set by the generator
LAST = {}                                      # records of the most recently constructed Model / MPC


class _Settings:
    def supress_ipopt_output(self):
        return None


class _Bounds(dict):
    def __setitem__(self, key, value):
        dict.__setitem__(self, tuple(key) if isinstance(key, tuple) else (key,), np.array(value, dtype=float))


class _Template(dict):
    """mpc.get_tvp_template(): records tvp_template['_tvp', :, name] = value."""

    def __setitem__(self, key, value):
        dict.__setitem__(self, key[-1], np.array(value, dtype=float))


class Model:
    def __init__(self, kind):
        self.kind = kind
        self.x, self.u, self.tvp, self.aux, self.rhs = {}, {}, {}, {}, {}
        LAST["model"] = self

    def set_variable(self, var_type, var_name, shape=(1, 1)):
        val = PROBE.get(var_type, {}).get(var_name)
        arr = np.zeros(shape) if val is None else np.array(val, dtype=float).reshape(shape)
        {"_x": self.x, "_u": self.u, "_tvp": self.tvp}[var_type][var_name] = arr
        return arr

    def set_rhs(self, name, expr):
        self.rhs[name] = np.array(expr, dtype=float)

    def set_expression(self, expr_name, expr):
        self.aux[expr_name] = np.array(expr, dtype=float)

    def setup(self):
        return None


class MPC:
    def __init__(self, model):
        self.model = model
        self.settings = _Settings()
        self.params, self.bounds, self.nl_cons = {}, _Bounds(), {}
        self.objective, self.rterm, self.tvp_fun = {}, {}, None
        self.rterm_calls = []
        self.x0 = None
        LAST["mpc"] = self

    def set_param(self, **kw):
        self.params.update(kw)

    def set_objective(self, mterm=None, lterm=None):
        self.objective = dict(mterm=np.array(mterm, dtype=float), lterm=np.array(lterm, dtype=float))

    def set_rterm(self, expr=None, **kw):
        """keyword form (mpc_cbf.py:180: set_rterm(u=R)) or expression form (optimal_decay_mpc_cbf.py:184-185, called twice):
        every call is recorded in order; what do-mpc does with repeated expression calls is NOT modelled here."""
        if expr is not None:
            self.rterm_calls.append(float(np.asarray(expr, dtype=float).reshape(-1)[0]))
        self.rterm.update({k: np.array(v, dtype=float) for k, v in kw.items()})

    def set_nl_cons(self, name, expr, ub=None, **kw):
        self.nl_cons[name] = (float(np.asarray(expr, dtype=float).reshape(-1)[0]), ub)

    def get_tvp_template(self):
        return _Template()

    def set_tvp_fun(self, f):
        self.tvp_fun = f

    def set_initial_guess(self):
        return None

    def setup(self):
        return None


class _Simulator:
    def __init__(self, model):
        self.model = model

    def set_param(self, **kw):
        return None

    def get_tvp_template(self):
        return _Template()

    def set_tvp_fun(self, f):
        return None

    def setup(self):
        return None


class _StateFeedback:
    def __init__(self, model):
        self.model = model


model = types.SimpleNamespace(Model=Model)
controller = types.SimpleNamespace(MPC=MPC)
simulator = types.SimpleNamespace(Simulator=_Simulator)
estimator = types.SimpleNamespace(StateFeedback=_StateFeedback)
graphics = types.SimpleNamespace(Graphics=lambda *a, **k: None)
