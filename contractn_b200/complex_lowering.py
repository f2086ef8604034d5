"""Complex operands on real kernels (the reference takes complex128 through numpy: ``abs`` gives a real scale,
contractn/einsum.py:97-106; every kernel of this engine is real).

A complex tensor ``X`` travels as a real tensor with one more, trailing index of extent 2 (re, im) -- for torch
that is ``view_as_real``, no copy.  A pairwise step ``x,y->z`` becomes the real three-operand contraction
``x p, p q r, y q -> z r`` with the 2x2x2 structure tensor of complex multiplication
(``S[0,0,0] = 1, S[1,1,0] = -1, S[0,1,1] = S[1,0,1] = 1``): first ``A~ S`` (a small re-arrangement of A), then ONE
real GEMM-class step whose summed extent is doubled -- the usual "complex GEMM as four real products", through the
same plan compiler and tensor-core kernels as any real step.

Every step is stabilised by the engine in ITS gauge (mean of |re| and |im| parts); the gauge of intermediate steps is
immaterial (``T = T_hat exp(c)`` holds exactly at every step), and the final result is re-normalised once to the
reference's canonical form, ``mean |T_hat| = 1`` with the complex modulus.
"""
from __future__ import annotations

import numpy as np

_POOL = "ΑΒΓΔΕΖΗΘΙΚΛΜΝΞΟΠΡΣΤΥΦΧΨΩ"


def _fresh(used, count):
    out = [ch for ch in _POOL if ch not in used]
    code = 0x4E00
    while len(out) < count:
        if chr(code) not in used:
            out.append(chr(code))
        code += 1
    return out[:count]


def structure_tensor(dtype):
    s = np.zeros((2, 2, 2), dtype=dtype)
    s[0, 0, 0] = 1.0
    s[1, 1, 0] = -1.0
    s[0, 1, 1] = 1.0
    s[1, 0, 1] = 1.0
    return s


def lower_step(einsum_str):
    """-> (real einsum string, explicit path, needs_structure_tensor) for one step of the pairwise list."""
    ins, out = einsum_str.split("->")
    ins = ins.split(",")
    p, q, r = _fresh(set(einsum_str), 3)
    if len(ins) == 1:
        return f"{ins[0]}{r}->{out}{r}", None, False
    if len(ins) == 2:
        return f"{ins[0]}{p},{p}{q}{r},{ins[1]}{q}->{out}{r}", ((0, 1), (0, 1)), True
    raise NotImplementedError("complex operands in a step with more than two inputs (%r)" % einsum_str)


def run(operands, contraction_list, contract_split, to_real, structure, to_complex, modulus_mean, log):
    """Walk the pairwise list; every step is one real contraction through ``contract_split(einstr, *ops,
    optimize=path) -> (T_hat, c)``.  The array-library specifics are injected:
    ``to_real(X) -> X~``, ``structure`` = the 2x2x2 tensor, ``to_complex(X~) -> X``,
    ``modulus_mean(X) -> (sum|X|, mean|X|)`` as python floats, ``log`` = math.log."""
    running = [to_real(op) for op in operands]
    log_scale = 0.0
    for step in contraction_list:
        inds, einsum_str = step[0], step[2]
        popped = [running.pop(x) for x in sorted(inds, reverse=True)]
        real_str, path, with_s = lower_step(einsum_str)
        ops = [popped[0], structure, popped[1]] if with_s else popped
        kwargs = {} if path is None else {"optimize": path}
        t, c = contract_split(real_str, *ops, **kwargs)
        log_scale += float(c)
        running.append(t)
    (result,) = running
    result = to_complex(result)
    total, mean = modulus_mean(result)
    if total > 1e-7:  # einsum.py:95-106 on the final tensor, complex modulus
        result = result / mean
        log_scale += log(mean)
    return result, log_scale
