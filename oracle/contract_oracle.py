"""numpy restatement of ContracTN's stabilised pairwise contraction executor.

TEST INFRASTRUCTURE -- never imported by the product path (contractn_b200/).

What it follows (all citations into /root/reference/contractn/):

* ``stabilize``        einsum.py:89-107   mean-abs rescale, 1e-7 skip on the SUM
* ``destabilize``      einsum.py:110-114
* ``core_contract``    einsum.py:326-393  pop order :344, routing :347,
                                           axes :357-368, result order :353-355,
                                           transpose :374-377, einsum :382-384
* ``contract``         einsum.py:190-310  always stabilise, then split or
                                           multiply the scale back

The per-step dispatch helpers the reference takes from the absent third-party
``opt_einsum`` (``_tensordot/_transpose/_einsum``) are thin wrappers over
``numpy.tensordot/transpose/einsum`` (the latter after remapping symbols onto
[a-zA-Z], which numpy requires); they are restated here from the published
opt_einsum 3.3.0 behaviour.

Parity pin: ``oracle/make_golden.py`` runs the UNMODIFIED reference package from
/root/reference in the build container (with ``oracle/oe_shim.py`` standing in
for the un-installable ``opt_einsum``) and commits its inputs/outputs under
``tests/golden/``; ``tests/test_oracle.py`` checks this restatement against
those vectors bit-for-bit, and against the README / notebook golden values
(README.md:33,56-57,74-76).  The contraction PATH itself is unpinned (no
reference test asserts one and opt_einsum's source is absent): see DESIGN.md.
"""
import operator
from functools import reduce

import numpy as np

MIN_NORM = 1e-7  # einsum.py:94
_AZ = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def stabilize(tensor, log_scale):
    """einsum.py:89-107 -- move the mean |entry| of ``tensor`` into ``log_scale``."""
    norm = np.sum(np.abs(tensor))
    numel = reduce(operator.mul, tensor.shape, 1)
    rescale = norm / numel
    nonneg = norm > MIN_NORM
    with np.errstate(divide="ignore", invalid="ignore"):
        tensor = np.where(nonneg, tensor / rescale, tensor)
        log_scale = np.where(nonneg, log_scale + np.log(rescale), log_scale)
    return tensor, log_scale


def destabilize(tensor, log_scale):
    """einsum.py:110-114."""
    with np.errstate(over="ignore"):
        return tensor * np.exp(log_scale)


def _einsum_az(einsum_str, *operands, **kwargs):
    """np.einsum after mapping arbitrary unicode symbols onto [a-zA-Z].

    Same relabelling rule as opt_einsum's ``_einsum`` (sorted symbols ->
    a, b, c, ...), so numpy sums in the same order as under the reference.
    """
    if all(ch in _AZ or ch in ",->" for ch in einsum_str):
        return np.einsum(einsum_str, *operands, **kwargs)
    symbols = sorted(set(einsum_str) - set(",->"))
    table = {s: _AZ[i] for i, s in enumerate(symbols)}
    return np.einsum("".join(table.get(ch, ch) for ch in einsum_str), *operands, **kwargs)


def core_contract(operands, contraction_list, **einsum_kwargs):
    """einsum.py:326-393 on numpy arrays.  Returns ``(tensor, log_scale)``."""
    work = list(operands)
    log_scale = np.zeros(())
    for positions, removed, step_str, _, blas_flag in contraction_list:
        taken = [work.pop(p) for p in positions]
        if blas_flag and "EINSUM" not in blas_flag:
            ins, wanted = step_str.split("->")
            left, right = ins.split(",")
            natural = "".join(s for s in left + right if s not in removed)
            if removed:
                pairs = sorted((left.find(s), right.find(s)) for s in removed)
                axes = tuple(zip(*pairs))
            else:
                axes = ((), ())
            new = np.tensordot(*taken, axes=axes)
            if natural != wanted:
                new = np.transpose(new, axes=tuple(map(natural.index, wanted)))
        else:
            new = _einsum_az(step_str, *taken, **einsum_kwargs)
        new, log_scale = stabilize(new, log_scale)
        work.append(new)
        del taken, new
    return work[0], log_scale


def contract(contraction_list, *operands, split_format=False, **einsum_kwargs):
    """einsum.py:303-310 given an already planned contraction list."""
    result, log_scale = core_contract(operands, contraction_list, **einsum_kwargs)
    if split_format:
        return result, log_scale
    return destabilize(result, log_scale)


def canonical_split(dense):
    """Finding 0.4: the split output is (T/mean|T|, ln mean|T|)."""
    dense = np.asarray(dense)
    kappa = np.sum(np.abs(dense)) / max(1, dense.size)
    return dense / kappa, np.log(kappa)
