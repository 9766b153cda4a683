"""`weight_function(w, sigma, qk)` handling.

The reference lets the user pass any Python callable (paint_with_words.py:402-405; README variants
with `qk.std()` and `log(1+sigma**2)`).  The fused kernel needs the bias in the factored form

    bias[n,t] = G(sigma) * stat(qk) * w[n,t]          stat in {max, unbiased std}

so the callable is *probed*, not executed on the score tensor (which is never materialised):
it is called with a recording stand-in for `qk` whose `.max()` / `.std()` return chosen scalars and a
tiny `w`; linearity in `w` and in the statistic is verified once per callable, and G(sigma) is then
read off per step by calling the user's own function -- so `0.4*w*math.log(1+sigma)*qk.max()`,
`0.5*w*math.log(1+sigma**2)*qk.std()`, and `lambda w, sigma, qk: 0.0` all work unchanged.
Callables outside this family raise `UnsupportedWeightFunction`: there is no fallback path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import torch

STAT_MAX, STAT_STD = 0, 1


class UnsupportedWeightFunction(TypeError):
    pass


class WeightFunction:
    """Explicit member of the family: `coef * w * log(1 + sigma**sigma_pow) * qk.<stat>()`.
    Also a valid `weight_function` for the reference itself (returns the same tensor)."""

    def __init__(self, coef: float = 0.1, sigma_pow: float = 1.0, stat: str = "max"):
        if stat not in ("max", "std"):
            raise ValueError("stat must be 'max' or 'std'")
        self.coef, self.sigma_pow, self.stat = float(coef), float(sigma_pow), stat

    def g(self, sigma) -> float:
        return self.coef * math.log(1.0 + float(sigma) ** self.sigma_pow)

    def __call__(self, w, sigma, qk):
        m = qk.max() if self.stat == "max" else qk.std()
        return self.coef * w * math.log(1 + float(sigma) ** self.sigma_pow) * m

    def __repr__(self):
        return f"WeightFunction(coef={self.coef}, sigma_pow={self.sigma_pow}, stat={self.stat!r})"


class _ScoreProbe:
    """Stand-in for the score tensor: only the two reductions the family uses are defined."""

    def __init__(self, vmax: float, vstd: float):
        self._vmax, self._vstd = torch.tensor(vmax), torch.tensor(vstd)
        self.used = set()

    def max(self):
        self.used.add("max")
        return self._vmax

    def std(self):
        self.used.add("std")
        return self._vstd


@dataclass(frozen=True)
class ProbedFunction:
    stat: Optional[int]           # STAT_MAX / STAT_STD, or None for an identically-zero function

    @property
    def is_zero(self) -> bool:
        return self.stat is None


_PROBE_CACHE: Dict[int, Tuple[Callable, ProbedFunction]] = {}


def _scalar(f, w_val: float, sigma, vmax: float, vstd: float) -> float:
    w = torch.full((1, 1), w_val, dtype=torch.float32)
    try:
        r = f(w, sigma, _ScoreProbe(vmax, vstd))
    except AttributeError as e:
        raise UnsupportedWeightFunction(
            f"weight_function uses an unsupported reduction of qk ({e}); only qk.max() and qk.std() "
            "can be fused") from e
    if isinstance(r, torch.Tensor):
        if r.numel() != 1:
            raise UnsupportedWeightFunction("weight_function must return a map shaped like w (or a scalar 0)")
        return float(r.reshape(()).item())
    return float(r)


def probe_weight_function(f: Callable, sigma=1.0) -> ProbedFunction:
    """Classify `f` (cached per callable object).  Raises UnsupportedWeightFunction."""
    hit = _PROBE_CACHE.get(id(f))
    if hit is not None and hit[0] is f:
        return hit[1]
    if isinstance(f, WeightFunction):
        res = ProbedFunction(STAT_MAX if f.stat == "max" else STAT_STD)
    else:
        s = sigma if float(sigma) > 0 else 1.0
        r_max = _scalar(f, 1.0, s, 1.0, 0.0)
        r_std = _scalar(f, 1.0, s, 0.0, 1.0)
        if r_max == 0.0 and r_std == 0.0:
            if _scalar(f, 1.0, s, 1.0, 1.0) != 0.0 or _scalar(f, 0.0, s, 1.0, 1.0) != 0.0:
                raise UnsupportedWeightFunction("weight_function is not linear in stat(qk)")
            res = ProbedFunction(None)
        elif r_max != 0.0 and r_std != 0.0:
            raise UnsupportedWeightFunction("weight_function mixes qk.max() and qk.std(); cannot be fused")
        else:
            use_max = r_max != 0.0
            base = r_max if use_max else r_std
            a, b = (3.0, 0.0) if use_max else (0.0, 3.0)
            lin_stat = _scalar(f, 1.0, s, a, b)
            lin_w = _scalar(f, 2.0, s, *((1.0, 0.0) if use_max else (0.0, 1.0)))
            zero_w = _scalar(f, 0.0, s, *((1.0, 0.0) if use_max else (0.0, 1.0)))
            tol = 1e-5 * abs(base)
            if abs(lin_stat - 3.0 * base) > 3 * tol or abs(lin_w - 2.0 * base) > 2 * tol or abs(zero_w) > tol:
                raise UnsupportedWeightFunction(
                    "weight_function is not of the form G(sigma) * w * stat(qk); cannot be fused")
            res = ProbedFunction(STAT_MAX if use_max else STAT_STD)
    while len(_PROBE_CACHE) >= 64:            # the reference-style loop builds a fresh uncond lambda every step: keep it bounded
        _PROBE_CACHE.pop(next(iter(_PROBE_CACHE)))
    _PROBE_CACHE[id(f)] = (f, res)
    return res


def g_of_sigma(f: Callable, probed: ProbedFunction, sigma) -> float:
    """G(sigma) = f(w=1, sigma, stat=1): the scalar that multiplies stat(qk) * w this step."""
    if probed.is_zero:
        return 0.0
    if isinstance(f, WeightFunction):
        return f.g(sigma)
    return _scalar(f, 1.0, sigma, 1.0, 0.0) if probed.stat == STAT_MAX else _scalar(f, 1.0, sigma, 0.0, 1.0)
