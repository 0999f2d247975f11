"""``B200GP``: the Gumbi ``Regressor`` backend whose dense path runs on the CUDA core.

Host-side mirror of ``gumbi.regression.pymc.GP.PymcGP`` (gumbi/regression/pymc/GP.py:21-979) for the hot path only:
``fit`` (:255-387), ``build_model`` (:468-583), ``_construct_kernels`` (:652-757), ``find_MAP`` (:799-813) and
``predict`` (:837-849) keep their names, arguments, ``MAP`` keys and error behaviour (``sparse=True`` selects the FITC
approximation of :571-578 on the device, ``gb2_fitc_*``), but no PyMC model is built: the
kernel composition is lowered to the plain ``spec`` dict of ``gumbi_b200._lib.make_kernel_struct`` and everything
O(N^2) and up happens behind the C ABI on the GPU.  There is no CPU fallback.

Two ways to use it:

* inside Gumbi (drop-in):  ``B200GP = gumbi_b200.make_backend(gumbi.regression.base.Regressor)`` gives a third
  ``Regressor`` subclass next to ``PymcGP``/``BotorchGP``; ``specify_model``, ``get_shaped_data``, ``prepare_grid``,
  ``predict_points``, ``predict_grid``, ``cross_validate`` ... are inherited unchanged.
* stand-alone on shaped arrays (what the tests on the GPU box do, where Gumbi is not installed):
  ``ArrayGP(X, y, continuous_dims=..., ...)`` -- a minimal stand-in for the state ``specify_model`` leaves behind.
"""
from __future__ import annotations

import numpy as np

from .engine import GPEngine

CONTINUOUS_KERNELS = ["ExpQuad", "Matern12", "Matern32", "Matern52", "Exponential"]
JITTER_DEFAULT = 1e-6  # pymc.gp.util.JITTER_DEFAULT


def assert_in(name, value, allowed):
    """Same message shape as gumbi.utils.misc.assert_in (used at pymc/GP.py:674)."""
    if value not in allowed:
        raise ValueError(f"{name} must be one of {allowed}, got {value!r}")


class B200Backend:
    """Mixin with the three backend methods + ``fit``.  Expects the attributes ``Regressor.specify_model`` sets."""

    # -- construction -----------------------------------------------------------------------------------------------
    def _init_backend(self, device=0, precision="fp64", distributed=False, multioutput="dense"):
        assert_in("multioutput", multioutput, ["dense", "kron", "auto"])
        self.device = device
        self.precision = precision
        self.distributed = distributed   # True: join torch.distributed's default group (one process per GPU)
        # "dense": factorise the stacked (nP x nP) system like the reference.  "kron": aligned multi-output models are split into
        # P independent n x n problems (gumbi_b200/kron.py; ValueError if the data are not aligned).  "auto": kron when possible.
        self.multioutput = multioutput
        self.engine = None
        self.MAP = None
        self.trace = None
        self.gp_dict = None
        self.model = None
        self.continuous_kernel = "ExpQuad"
        self.heteroskedastic_inputs = False
        self.heteroskedastic_outputs = True
        self.sparse = False
        self.latent = False
        self.n_u = 100
        self.ARD = True
        self._factor_key = None
        # lengthscale prior of find_MAP: "InverseGamma" (today's reference, GP.py:385,407) or "Gamma(2,1)" (its older code, GP.py:408)
        self.ls_prior = "InverseGamma"
        self.model_specs = {
            "seed": self.seed,
            "continuous_kernel": self.continuous_kernel,
            "heteroskedastic_inputs": self.heteroskedastic_inputs,
            "heteroskedastic_outputs": self.heteroskedastic_outputs,
            "sparse": self.sparse,
            "n_u": self.n_u,
        }

    # -- fit (GP.py:255-387) ------------------------------------------------------------------------------------------
    def fit(self, outputs=None, linear_dims=None, continuous_dims=None, continuous_levels=None, continuous_coords=None,
            categorical_dims=None, categorical_levels=None, additive=False, seed=None, continuous_kernel="ExpQuad",
            period=None, heteroskedastic_inputs=False, heteroskedastic_outputs=True, sparse=False, n_u=100, ARD=True,
            ls_bounds=None, mass=0.98, spec_kwargs=None, build_kwargs=None, MAP_kwargs=None):
        self.specify_model(outputs=outputs, linear_dims=linear_dims, continuous_dims=continuous_dims,
                           continuous_levels=continuous_levels, continuous_coords=continuous_coords,
                           categorical_dims=categorical_dims, categorical_levels=categorical_levels, additive=additive,
                           **(spec_kwargs or {}))
        self.build_model(seed=seed, continuous_kernel=continuous_kernel, period=period,
                         heteroskedastic_inputs=heteroskedastic_inputs, heteroskedastic_outputs=heteroskedastic_outputs,
                         sparse=sparse, n_u=n_u, ARD=ARD, ls_bounds=ls_bounds, mass=mass, **(build_kwargs or {}))
        self.find_MAP(**(MAP_kwargs or {}))
        return self

    # -- build_model (GP.py:468-583) ----------------------------------------------------------------------------------
    def build_model(self, seed=None, continuous_kernel="ExpQuad", period=None, heteroskedastic_inputs=False,
                    heteroskedastic_outputs=True, sparse=False, n_u=100, ARD=True, ls_bounds=None, mass=0.98):
        if heteroskedastic_inputs:
            raise NotImplementedError("Heteroskedasticity over inputs is not yet implemented.")
        if sparse and (self.distributed or self.precision != "fp64"):
            raise NotImplementedError("sparse=True (FITC) runs on one GPU in fp64")
        kernels = CONTINUOUS_KERNELS + ["Periodic"]
        kernels += [k + "+Periodic" for k in kernels if k != "Periodic"]          # GP.py:664-674
        assert_in("Continuous kernel", continuous_kernel, kernels)
        if "Periodic" in continuous_kernel and period is None:
            raise ValueError("Period must be specified for periodic kernel")     # GP.py:678-679

        X, y = self.get_shaped_data("mean")
        D_in = len(self.dims)
        assert X.shape[1] == D_in

        seed = self.seed if seed is None else seed
        self.seed = seed
        self.continuous_kernel = continuous_kernel
        self.heteroskedastic_inputs = heteroskedastic_inputs
        self.heteroskedastic_outputs = heteroskedastic_outputs
        self.sparse = sparse
        self.n_u = n_u
        self.latent = False
        self.ARD = ARD
        self.ls_bounds = ls_bounds
        self.mass = mass
        self.period = period
        self.model_specs = {
            "seed": seed,
            "continuous_kernel": continuous_kernel,
            "heteroskedastic_inputs": heteroskedastic_inputs,
            "heteroskedastic_outputs": heteroskedastic_outputs,
            "sparse": sparse,
            "n_u": n_u,
        }
        self._X = np.ascontiguousarray(X, dtype=np.float64)
        self._y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        self._layout = self._model_layout()
        self._layout["warp"] = self._periodic_warp(continuous_kernel, period)
        self._Xu = None
        if sparse:
            # GP.py:571-578: k-means inducing points over the full shaped X, MarginalSparse(approx="FITC") with the scalar sigma
            import warnings

            from .sparse import kmeans_inducing_points

            if heteroskedastic_outputs and self._layout["multi"]:
                warnings.warn("Heteroskedasticity over outputs is not yet implemented for sparse GP. Reverting to scalar-valued noise.")
            self._Xu = kmeans_inducing_points(n_u, self._X, seed=seed)
        want_kron = (not sparse) and self._wants_kron()
        if self.engine is not None and want_kron != (type(self.engine).__name__ == "KronEngine"):
            self.engine.close()      # the structure changed between two build_model calls
            self.engine = None
        if self.engine is None:
            self.engine = self._make_kron_engine() if want_kron else self._make_dense_engine()
        self.engine.set_train(self._engine_points(self._X), self._y)
        self._factor_key = None
        self.model = self._layout  # truthy placeholder: the reference asserts ``self.model is not None`` in find_MAP
        return self

    # -- periodic kernels (GP.py:389-447) ---------------------------------------------------------------------------
    def _periodic_warp(self, continuous_kernel, period):
        """How ``continuous_kernel="Periodic"`` / ``"<K>+Periodic"`` are lowered onto the stationary kernels of the CUDA core.

        Both are stationary kernels of the warped coordinates u(x) = [sin(2 pi x / T), cos(2 pi x / T)] (the reference's own
        ``mapping`` for ``WarpedInput``, GP.py:432-435), because |u(x) - u(x')|^2 = 2 - 2 cos(2 pi (x - x') / T) = 4 sin^2(pi (x - x') / T):
          * ``pm.gp.cov.Periodic``: exp(-1/2 sum_j sin^2(pi (x_j - x'_j) / T_j) / ls_j^2) = ExpQuad on u with lengthscale 2 ls_j;
          * ``"<K>+Periodic"``: ``WarpedInput(cov_func=K(input_dim=2, ls), warp_func=mapping)`` = K on u with lengthscale ls.
        The warped columns are appended to every array handed to the engine (O(N) host work); the Linear kernel and the Coregion
        factors keep reading the original columns.  Returns None for the plain kernels."""
        if "Periodic" not in continuous_kernel:
            return None
        lay = self._layout
        if hasattr(period, "z"):   # ParameterArray: the reference reads period.z[dim + "_z"] (GP.py:403, :429)
            zp = [np.asarray(period.z[dim + "_z"].values(), dtype=np.float64).reshape(-1)[0] for dim in self.continuous_dims]
        else:                      # plain arrays (ArrayGP): {dim: standardized period}
            zp = [float(period[dim]) for dim in self.continuous_dims]
        zp = np.asarray(zp, dtype=np.float64)
        if not np.all(np.isfinite(zp)) or np.any(zp == 0):
            raise ValueError("periods must be finite and non-zero")
        base = continuous_kernel.removesuffix("+Periodic")
        if continuous_kernel != "Periodic" and lay["n_s"] != 1:
            # the inner kernel is built with input_dim=2 (GP.py:424-426): with several continuous dims it would only see the first
            # two warped columns -- not a model anybody means
            raise NotImplementedError('"<kernel>+Periodic" is defined for exactly one continuous dimension')
        D = len(self.dims)
        return {"mode": "periodic" if continuous_kernel == "Periodic" else "warped", "kind": "ExpQuad" if continuous_kernel == "Periodic" else base,
                "c": 2.0 * np.pi / zp, "src": list(lay["idx_s"]), "cols": [D + j for j in range(2 * lay["n_s"])]}

    def _engine_points(self, points):
        """Columns the engine sees: the model dims, plus the warped coordinates of a periodic kernel."""
        w = self._layout.get("warp")
        if not w:
            return points
        x = points[:, w["src"]] * w["c"][None, :]
        return np.ascontiguousarray(np.hstack([points, np.sin(x), np.cos(x)]))

    def _engine_ls(self, ls):
        """Lengthscales of the engine's stationary kernel for the model's ``ls`` (see ``_periodic_warp``)."""
        w = self._layout.get("warp")
        ls = np.atleast_1d(np.asarray(ls, dtype=np.float64))
        if not w:
            return ls
        if w["mode"] == "warped":
            return ls.reshape(-1)[:1]                                   # one lengthscale shared by sin and cos
        return 2.0 * (np.concatenate([ls, ls]) if ls.size == self._layout["n_s"] and ls.size > 1 else ls.reshape(-1)[:1])

    def _fold_ls_gradient(self, g):
        """Gradient w.r.t. the engine's lengthscales -> gradient w.r.t. the model's ``ls``."""
        w = self._layout.get("warp")
        g = np.atleast_1d(np.asarray(g, dtype=np.float64))
        if not w:
            return g
        if w["mode"] == "warped":
            return np.array([g.sum()])
        n_s = self._layout["n_s"]
        if g.size == 2 * n_s and n_s > 1 and self.ARD:
            return 2.0 * (g[:n_s] + g[n_s:])
        return np.array([2.0 * g.sum()])

    def _make_dense_engine(self):
        engine = GPEngine(self.device, self.precision)
        if self.distributed:
            from . import dist as gdist

            gdist.init_engine(engine)   # collective: factorize() is row-block sharded across the ranks from here on
        return engine

    def _make_block_engine(self):
        """One single-GPU engine per output block of the Kronecker-aware solve (never joined to the process group)."""
        return GPEngine(self.device, self.precision)

    def _make_kron_engine(self):
        from . import kron

        rank, world, gather = 0, 1, None
        if self.distributed:
            import torch.distributed as tdist

            if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1:
                rank, world = tdist.get_rank(), tdist.get_world_size()

                def gather(obj):
                    parts = [None] * world
                    tdist.all_gather_object(parts, obj)
                    return parts

        lay = self._layout
        pcol, P = lay["terms"][0]["coreg"][-1][1], lay["terms"][0]["coreg"][-1][2]
        return kron.KronEngine(self._make_block_engine, pcol, P, rank=rank, world=world, gather=gather)

    def _wants_kron(self):
        """Whether this model is solved block-wise (see ``multioutput``): needs the output Coregion, a non-additive model and
        aligned observations (every output at the same inputs)."""
        if self.multioutput == "dense":
            return False
        from . import kron

        try:
            if self.out_col not in self.categorical_dims:
                raise kron.NotAligned("the model has a single output")
            if self.additive:
                raise kron.NotAligned("additive models have no single B (x) Kx form")
            pcol = self._get_dim_indexes()["p"]
            kron.aligned_blocks(self._X, pcol, len(self.categorical_levels[self.out_col]))
        except kron.NotAligned as e:
            if self.multioutput == "kron":
                raise ValueError(f'multioutput="kron" is not applicable: {e}') from None
            return False
        return True

    # -- structure of the covariance (GP.py:652-757) ------------------------------------------------------------------
    def _get_dim_counts(self):
        return {"l": len(self.linear_dims), "s": len(self.continuous_dims), "c": len(self.categorical_dims),
                "p": len(self.outputs)}

    def _get_dim_indexes(self):
        return {
            "l": [self.dims.index(dim) for dim in self.linear_dims],
            "s": [self.dims.index(dim) for dim in self.continuous_dims],
            "c": [self.dims.index(dim) for dim in self.categorical_dims],
            "p": self.dims.index(self.out_col) if self.out_col in self.dims else None,
        }

    def _model_layout(self):
        """Which random variables exist and how they combine -- the structure ``_construct_kernels`` builds.

        Returns a list of terms; each term is ``{"suffix", "coreg": [(name, col, P), ...]}`` where ``name`` keys the
        ``W_<name>``/``kappa`` entries of MAP.  The output Coregion (``W_<out_col>``) is shared by all terms (:724-727,:749).
        """
        ns, idxs = self._get_dim_counts(), self._get_dim_indexes()
        multi = self.out_col in self.categorical_dims
        out_cg = (self.out_col, idxs["p"], len(self.categorical_levels[self.out_col])) if multi else None
        terms = []
        total = {"suffix": "total", "coreg": []}
        if ns["c"] > 0 and not self.additive:
            for dim, idx in zip(self.categorical_dims, idxs["c"]):
                if dim == self.out_col:
                    continue
                total["coreg"].append((dim, idx, len(self.categorical_levels[dim])))
        if multi:
            total["coreg"].append(out_cg)
        terms.append(total)
        if self.additive:
            for dim, idx in zip(self.categorical_dims, idxs["c"]):
                if dim == self.out_col:
                    continue
                t = {"suffix": dim, "coreg": [(dim, idx, len(self.categorical_levels[dim]))]}
                if multi:
                    t["coreg"].append(out_cg)
                terms.append(t)
        noise_cg = None
        if self.heteroskedastic_outputs and multi and not self.sparse:   # sparse: scalar sigma (GP.py:573-578)
            noise_cg = ("Output_noise", idxs["p"], out_cg[2])
        return {"terms": terms, "noise_coreg": noise_cg, "idx_s": idxs["s"], "idx_l": idxs["l"], "n_s": ns["s"],
                "n_l": ns["l"], "multi": multi}

    def param_shapes(self):
        """Names and shapes of the free hyper-parameters, in the order PyMC would register them."""
        lay = self._layout
        shapes = {}
        seen = set()
        for t in lay["terms"]:
            sfx = t["suffix"]
            shapes[f"ls_{sfx}"] = (lay["n_s"] if self.ARD else 1,)
            shapes[f"η_{sfx}"] = ()
            if lay["n_l"] > 0:
                shapes[f"c_{sfx}"] = (lay["n_l"],)
                shapes[f"τ_{sfx}"] = ()
            for name, _, P in t["coreg"]:
                if name not in seen:
                    seen.add(name)
                    shapes[f"W_{name}"] = (P, 2)
                    shapes[f"κ_{name}"] = (P,)
        shapes["σ"] = ()
        if lay["noise_coreg"]:
            name, _, P = lay["noise_coreg"]
            shapes[f"W_{name}"] = (P, 2)
            shapes[f"κ_{name}"] = (P,)
        return shapes

    def spec_from_point(self, point):
        """MAP-style dict of hyper-parameter values -> kernel ``spec`` for the CUDA core."""
        lay = self._layout
        terms = []
        for t in lay["terms"]:
            sfx = t["suffix"]
            warp = lay.get("warp")
            ls = self._engine_ls(point[f"ls_{sfx}"])
            term = {"kind": warp["kind"] if warp else self.continuous_kernel, "cont_idx": list(warp["cols"] if warp else lay["idx_s"]), "ls": ls.tolist(),
                    "eta": float(point[f"η_{sfx}"]), "lin_idx": [], "c": [], "tau": 0.0, "coreg": []}
            if lay["n_l"] > 0:
                term["lin_idx"] = list(lay["idx_l"])
                term["c"] = np.atleast_1d(np.asarray(point[f"c_{sfx}"], dtype=np.float64)).tolist()
                term["tau"] = float(point[f"τ_{sfx}"])
            for name, col, _ in t["coreg"]:
                term["coreg"].append({"col": col, "W": np.asarray(point[f"W_{name}"]).tolist(),
                                      "kappa": np.asarray(point[f"κ_{name}"]).tolist()})
            terms.append(term)
        spec = {"terms": terms, "sigma": float(point["σ"]), "noise_coreg": None, "jitter": JITTER_DEFAULT}
        if lay["noise_coreg"]:
            name, col, _ = lay["noise_coreg"]
            spec["noise_coreg"] = {"col": col, "W": np.asarray(point[f"W_{name}"]).tolist(),
                                   "kappa": np.asarray(point[f"κ_{name}"]).tolist()}
        return spec

    # -- find_MAP (GP.py:799-813) -------------------------------------------------------------------------------------
    def find_MAP(self, *args, **kwargs):
        """Maximum a posteriori hyper-parameters (``pm.find_MAP``: L-BFGS-B on -(logp + log-priors), jacobian=False).

        ``start=`` may carry initial values; ``point=`` pins the hyper-parameters outright (no optimisation), which is
        how a maintainer feeds an existing PyMC ``MAP`` dict to this backend.
        """
        assert self.model is not None
        point = kwargs.pop("point", None)
        if point is not None:
            self.MAP = self._complete_point(point)
        else:
            from .map import find_map  # local import: scipy only needed for fitting

            self.MAP = find_map(self, *args, **kwargs)
        self._factor_key = None
        return self.MAP

    def _complete_point(self, point):
        shapes = self.param_shapes()
        out = {}
        for name, shape in shapes.items():
            if name not in point:
                raise KeyError(f"hyper-parameter {name!r} missing from point (expected {sorted(shapes)})")
            val = np.asarray(point[name], dtype=np.float64)
            if name.startswith("ls_") and val.size == 1:
                val = val.reshape(1) if shape == (1,) else np.repeat(val.reshape(1), shape[0])
            if val.shape != tuple(shape):
                raise ValueError(f"hyper-parameter {name!r} has shape {val.shape}, expected {tuple(shape)}")
            out[name] = val
            if name.split("_")[0] in ("ls", "η", "τ", "κ", "σ"):
                if np.any(val <= 0):
                    raise ValueError(f"hyper-parameter {name!r} must be positive")
                out[name + "_log__"] = np.log(val)  # pm.find_MAP also returns the transformed values
        return out

    # -- inference modes that stay with PyMC (GP.py:759-797, :815-835) -------------------------------------------------
    def build_latent(self, *args, **kwargs):
        raise NotImplementedError("Latent GPs (pm.gp.Latent) are out of scope of the B200 backend; use PymcGP.build_latent.")

    def sample(self, *args, **kwargs):
        raise NotImplementedError("MCMC over the hyper-parameters is out of scope of the B200 backend (MAP inference only); use PymcGP.sample.")

    # -- predict (GP.py:837-849) --------------------------------------------------------------------------------------
    def _ensure_factorized(self):
        if self.MAP is None:
            raise RuntimeError("predict called before find_MAP/fit")
        # fingerprint of the hyper-parameter VALUES (a MAP dict edited in place must trigger a new factorisation)
        key = self._map_key()
        if self._factor_key != key:
            self.engine.set_kernel(self.spec_from_point(self.MAP))
            if self.sparse:
                self.engine.fitc_factorize(self._engine_points(self._Xu))
            else:
                self.engine.factorize()
            self._factor_key = key

    def _map_key(self):
        return hash(tuple((k, np.asarray(v, dtype=np.float64).tobytes()) for k, v in sorted(self.MAP.items()) if not k.endswith("_log__")))

    def _ensure_factorized_key(self):
        """Record that the engine now holds the factor of the current MAP (after a fused factorise+predict)."""
        self._factor_key = self._map_key()

    def predict(self, points_array, with_noise=True, additive_level="total", **kwargs):
        if additive_level != "total":
            raise NotImplementedError("Prediction for additive sublevels is not yet supported.")
        if kwargs:
            raise TypeError(f"unsupported predict arguments for the B200 backend: {sorted(kwargs)}")
        self._ensure_factorized()
        points_array = np.atleast_2d(np.asarray(points_array, dtype=np.float64))
        if self.sparse:
            return self.engine.fitc_predict(self._engine_points(points_array), pred_noise=bool(with_noise))
        if getattr(self.engine, "world", 1) > 1 and not getattr(self.engine, "shard_storage", False):
            # every rank holds the factor: each serves a contiguous slice of the points, slices are gathered on all ranks
            from . import dist as gdist

            M, world = len(points_array), self.engine.world
            lo, hi = gdist.grid_slice(M, self.engine.rank, world)
            if hasattr(self.engine, "allgather_device"):
                # slices stay on the device: predict into [mean | var] of a padded slot, one NCCL all-gather on the handle's
                # stream (gb2_dist_allgather_dev), one D2H copy -- no pickling of host arrays through torch.distributed
                import torch

                dev = torch.device("cuda", self.engine.device)
                slot = -(-M // world)
                dXs = torch.from_numpy(np.ascontiguousarray(self._engine_points(points_array[lo:hi]))).to(dev)
                dloc = torch.zeros(2 * slot, dtype=torch.float64, device=dev)
                dall = torch.empty(2 * slot * world, dtype=torch.float64, device=dev)
                torch.cuda.synchronize(dev)
                if hi > lo:
                    self.engine.predict_device(dXs.data_ptr(), hi - lo, bool(with_noise), dloc.data_ptr(), dloc.data_ptr() + 8 * slot)
                self.engine.allgather_device(dloc.data_ptr(), dall.data_ptr(), 2 * slot)
                a = dall.cpu().numpy().reshape(world, 2, slot)
                cnt = [gdist.grid_slice(M, r, world)[1] - gdist.grid_slice(M, r, world)[0] for r in range(world)]
                return (np.concatenate([a[r, 0, :cnt[r]] for r in range(world)]), np.concatenate([a[r, 1, :cnt[r]] for r in range(world)]))
            mu, var = self.engine.predict(self._engine_points(points_array[lo:hi]), pred_noise=bool(with_noise))
            return gdist.gather_grid(mu, var, M)
        return self.engine.predict(self._engine_points(points_array), pred_noise=bool(with_noise))

    # -- conditional / posterior samples (GP.py:861-979) ------------------------------------------------------------
    def conditional(self, points_array, pred_noise=False):
        """Mean and full covariance of ``gp_dict["total"].conditional(var_name, points_array)`` (GP.py:913-914), on device."""
        if self.sparse:
            raise NotImplementedError("full-covariance conditionals of the sparse (FITC) model are not implemented")
        self._ensure_factorized()
        points_array = np.atleast_2d(np.asarray(points_array, dtype=np.float64))
        return self.engine.predict_full(self._engine_points(points_array), pred_noise=bool(pred_noise))

    def sample_conditional(self, points_array, size=1, random_seed=None, pred_noise=False):
        """``size`` joint draws from the conditional at ``points_array`` (standardized space), shape (size, M).

        The O(N^2 M + N M^2) part (solve, covariance) runs on the GPU; the draw itself is mu + chol(cov + 1e-6 I) z on the host
        (PyMC's ``stabilize`` jitter), O(M^3) for the M requested points only."""
        mu, cov = self.conditional(points_array, pred_noise=pred_noise)
        rng = np.random.default_rng(self.seed if random_seed is None else random_seed)
        Lc = np.linalg.cholesky(cov + JITTER_DEFAULT * np.eye(len(mu)))
        return mu[None, :] + rng.standard_normal((int(size), len(mu))) @ Lc.T

    def draw_point_samples(self, points, *args, source=None, output=None, var_name="posterior_samples", additive_level="total",
                           increment_var=True, size=1, random_seed=None, **kwargs):
        """Mirror of ``PymcGP.draw_point_samples`` (GP.py:861-922) for a MAP source: joint posterior draws at ``points``.

        Needs the ``Regressor`` base class (``_parse_prediction_output``, ``_prepare_points_for_prediction``, ``parray``)."""
        if additive_level != "total":
            raise NotImplementedError("Prediction for additive sublevels is not yet supported.")
        output = self._parse_prediction_output(output)
        if len(output) > 1:
            raise NotImplementedError("Drawing correlated samples of multiple outputs is not yet implemented.")
        points_array, tall_points, param_coords = self._prepare_points_for_prediction(points, output=output)
        if source is None:
            if self.MAP is None:
                raise ValueError('"Source" of predictions must be supplied if GP object has no trace or MAP stored.')
        elif isinstance(source, dict):
            self.find_MAP(point=source)
        else:
            raise NotImplementedError("The B200 backend draws posterior samples from a MAP point only (no MCMC trace).")
        samples = self.sample_conditional(points_array, size=size, random_seed=random_seed)
        self.predictions = self.parray(**{output[0]: samples}, stdzd=True)
        self.predictions_X = points
        return self.predictions

    def draw_grid_samples(self, *args, source=None, output=None, categorical_levels=None, var_name="posterior_samples",
                          additive_level="total", increment_var=True, **kwargs):
        """Mirror of ``PymcGP.draw_grid_samples`` (GP.py:924-979)."""
        if self.grid_points is None:
            raise ValueError("Grid must first be specified with `prepare_grid`")
        points = self.grid_points
        if self.categorical_dims:
            points = self.append_categorical_points(points, categorical_levels=categorical_levels)
        samples = self.draw_point_samples(*args, points=points, output=output, source=source, var_name=var_name,
                                          additive_level=additive_level, increment_var=increment_var, **kwargs)
        self.predictions = samples.reshape(-1, *self.grid_parray.shape)
        self.predictions_X = self.predictions_X.reshape(self.grid_parray.shape)
        return self.predictions

    def predict_cold(self, points_array, with_noise=True, fused=False):
        """What ONE reference ``predict`` call costs: rebuild K, re-factorise, solve.  For benchmarking.

        ``fused=True``: one pass through ``gb2_factorize_predict`` (prediction points carried through the factorisation as extra
        rows of the factor) instead of factorise-then-solve; single GPU, fp64, dense solver only."""
        self._factor_key = None
        if not fused or self.sparse:
            return self.predict(points_array, with_noise=with_noise)
        if self.MAP is None:
            raise RuntimeError("predict called before find_MAP/fit")
        if not hasattr(self.engine, "factorize_predict") or getattr(self.engine, "world", 1) > 1:
            raise NotImplementedError("fused cold predict needs the dense single-GPU engine")
        points_array = np.atleast_2d(np.asarray(points_array, dtype=np.float64))
        self.engine.set_kernel(self.spec_from_point(self.MAP))
        out = self.engine.factorize_predict(self._engine_points(points_array), pred_noise=bool(with_noise))
        self._ensure_factorized_key()
        return out

    def marginal_log_likelihood(self, point=None):
        """log p(y | X, theta) of ``gp.marginal_likelihood("ml", ...)`` (GP.py:580) at ``point`` (default: MAP)."""
        if point is not None:
            self.engine.set_kernel(self.spec_from_point(self._complete_point(point)))
            if self.sparse:
                self.engine.fitc_factorize(self._engine_points(self._Xu))
            else:
                self.engine.factorize()
            self._factor_key = None
        else:
            self._ensure_factorized()
        return self.engine.fitc_mll() if self.sparse else self.engine.mll()


class ArrayRegressor:
    """Minimal stand-in for the state a ``gumbi.regression.base.Regressor`` holds after ``specify_model``.

    It carries already-shaped arrays (what ``get_shaped_data`` returns, base.py:435-471) so that the backend can be
    exercised where Gumbi itself is not installed (the GPU box).  Columns of ``X`` follow ``self.dims`` =
    continuous_dims + categorical_dims, with the output column last (base.py:155-158).
    """

    def __init__(self, X, y, continuous_dims, linear_dims=None, categorical_dims=None, categorical_levels=None,
                 out_col="Variable", outputs=None, additive=False, seed=2021):
        self._X_shaped = np.atleast_2d(np.asarray(X, dtype=np.float64))
        self._y_shaped = np.asarray(y, dtype=np.float64).reshape(-1)
        self.continuous_dims = list(continuous_dims)
        self.linear_dims = list(linear_dims or [])
        self.categorical_dims = list(categorical_dims or [])
        self.categorical_levels = dict(categorical_levels or {})
        self.out_col = out_col
        self.outputs = list(outputs or ["y"])
        self.additive = additive
        self.seed = seed
        self.filter_dims = {}
        self.model_specs = {}
        if self._X_shaped.shape[1] != len(self.dims):
            raise ValueError(f"X has {self._X_shaped.shape[1]} columns but dims = {self.dims}")
        for dim in self.linear_dims:
            if dim not in self.continuous_dims:
                raise ValueError("linear_dims must be a subset of continuous_dims")

    @property
    def dims(self):
        return self.continuous_dims + self.categorical_dims

    def specify_model(self, **kwargs):
        given = {k: v for k, v in kwargs.items() if v not in (None, False)}
        if given:
            raise TypeError(f"ArrayRegressor is specified at construction; got {sorted(given)}")

    def get_shaped_data(self, metric="mean", dropna=True):
        nans = np.isnan(self._y_shaped)
        return self._X_shaped[~nans], self._y_shaped[~nans]


class ArrayGP(B200Backend, ArrayRegressor):
    def __init__(self, X, y, continuous_dims, device=0, precision="fp64", distributed=False, multioutput="dense", **kwargs):
        ArrayRegressor.__init__(self, X, y, continuous_dims, **kwargs)
        self._init_backend(device=device, precision=precision, distributed=distributed, multioutput=multioutput)


def make_backend(regressor_base, device=0, precision="fp64", distributed=False, multioutput="dense"):
    """Create the drop-in ``B200GP`` subclass of Gumbi's ``Regressor`` (``gumbi.regression.base.Regressor``).

    The keyword arguments become the class's constructor defaults, so that objects Gumbi re-creates itself with
    ``self.__class__(dataset, outputs=..., seed=...)`` (``cross_validate``, base.py:1060) keep the device / precision / solver."""
    defaults = dict(device=device, precision=precision, distributed=distributed, multioutput=multioutput)

    class B200GP(B200Backend, regressor_base):
        def __init__(self, dataset, outputs=None, seed=2021, **options):
            unknown = set(options) - set(defaults)
            if unknown:
                raise TypeError(f"unexpected keyword arguments {sorted(unknown)} (options: {sorted(defaults)})")
            regressor_base.__init__(self, dataset, outputs, seed)
            self._init_backend(**{**defaults, **options})

    B200GP.__doc__ = "Gumbi Regressor backend running the exact-GP dense path on a B200 (see gumbi_b200.backend)."
    return B200GP


def __getattr__(name):
    if name == "B200GP":  # resolved lazily so that importing gumbi_b200 never requires gumbi
        try:
            from gumbi.regression.base import Regressor
        except Exception as e:  # pragma: no cover - depends on the environment
            raise ImportError("gumbi is not importable; use gumbi_b200.ArrayGP or make_backend(Regressor)") from e
        cls = make_backend(Regressor)
        globals()["B200GP"] = cls
        return cls
    raise AttributeError(name)

