"""Reverse mode for the device plan.

The reference's torch backend is plain differentiable torch ops (``contractn/einsum.py:326-393``:
``_core_contract`` loops ``tensordot`` / ``einsum`` and ``stabilize``, ``einsum.py:89-107``), so a
``TN.contract()`` on tensors that require gradients can be back-propagated through.  The device
plan is one opaque sequence of kernels; its backward is built here from what it computes:

* forward: ``T = einsum(X_1 .. X_N)``, returned either dense or in split format
  ``(T_hat, c) = (T / kappa, ln kappa)``, ``kappa = mean|T|`` (the canonical form every plan ends in);
* a contraction is linear in each operand, so ``dL/dX_k`` is itself a contraction: of ``G = dL/dT``
  with all the OTHER operands, output labels = the labels of ``X_k`` -- one more plan per operand
  that asks for a gradient, run by the same engine, stabilised like any other contraction;
* split format: with upstream gradients ``(g, g_c)`` for ``(T_hat, c)``,
  ``dL/dT_j = exp(-c) * G_hat_j``, ``G_hat_j = g_j + sign(T_hat_j)/n * (g_c - sum_i g_i T_hat_i)``
  (chain rule through ``T_hat = T/kappa``, ``c = ln kappa``, ``d kappa = mean(sign(T) dT)``).
  ``G_hat`` is O(1) where ``T`` itself may be far outside the float range; the factor ``exp(-c)`` is
  applied to the LOG-SCALE of the gradient contraction, never to data.

Second derivatives are not provided (the backward runs under ``no_grad``).
"""
from __future__ import annotations

__all__ = ["global_einsum", "gradient_einsum", "split_output_cotangent"]


def _replay(contraction_list, n_ops):
    """-> (list of input label strings, output label string)."""
    # running list of (original operand index or None, labels or None)
    running = [(i, None) for i in range(n_ops)]
    inputs = [None] * n_ops
    out = None
    for step in contraction_list:
        inds, einsum_str = step[0], step[2]
        ins, out = einsum_str.split("->")
        ins = ins.split(",")
        popped = [running.pop(x) for x in sorted(inds, reverse=True)]
        if len(popped) != len(ins):
            raise ValueError("malformed contraction list")
        for (orig, _), lab in zip(popped, ins):
            if orig is not None:
                inputs[orig] = lab
        running.append((None, out))
    if any(lab is None for lab in inputs):
        raise ValueError("contraction list does not consume every operand")
    return inputs, out


def global_einsum(contraction_list, n_ops):
    """``(input_labels, output_labels)`` of the whole contraction, rebuilt from its pairwise list: a step
    is ``(positions, removed, 'x,y->z', remaining, kind)`` whose operands were popped from the running
    operand list at ``positions`` in descending order (``einsum.py:347``) and whose result was appended."""
    return _replay(list(contraction_list), n_ops)


def gradient_einsum(inputs, output, k, shapes):
    """The contraction that yields ``dL/dX_k`` from ``G = dL/dT`` and the other operands.

    Returns ``(einsum_string, operand_spec)``; ``operand_spec`` lists what to pass, in order:
    ``("G",)``, ``("X", m)`` for operand ``m``, or ``("ones", dim)`` for a label of ``X_k`` that occurs
    nowhere else (a plain sum over that index in the forward: its gradient is a broadcast).
    A scalar output contributes no operand (``G`` is then a number the caller multiplies in).
    """
    target = inputs[k]
    if len(set(target)) != len(target):
        raise NotImplementedError("gradient w.r.t. an operand with a repeated label (%r)" % target)
    terms, spec = [], []
    if output:
        terms.append(output)
        spec.append(("G",))
    for m, lab in enumerate(inputs):
        if m != k:
            terms.append(lab)
            spec.append(("X", m))
    seen = set("".join(terms))
    for pos, ch in enumerate(target):
        if ch not in seen:
            terms.append(ch)
            spec.append(("ones", shapes[k][pos]))
    return ",".join(terms) + "->" + target, spec


def split_output_cotangent(t_hat, g_t, g_c):
    """``G_hat`` with ``dL/dT = exp(-c) * G_hat`` for a split-format result (module docstring).  torch tensors."""
    import torch

    n = t_hat.numel()
    g_t = torch.zeros_like(t_hat) if g_t is None else g_t.to(t_hat.dtype)
    inner = (g_t * t_hat).sum()
    g_c = inner.new_zeros(()) if g_c is None else g_c.to(inner.dtype)
    return g_t + torch.sign(t_hat) * ((g_c - inner) / n)


def make_function(forward_nograd, contract_split):
    """Build the ``torch.autograd.Function`` over two callables (injected so that the algebra can be
    tested on the CPU against torch's own autograd):

    ``forward_nograd(tensors, contraction_list, out_mode) -> (out, log_scale)`` and
    ``contract_split(einsum_string, *operands) -> (R_hat, c_r)``.
    """
    import torch

    class ContractFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, contraction_list, out_mode, dense, *tensors):
            out, log_scale = forward_nograd([t.detach() for t in tensors], contraction_list, out_mode)
            ctx.contraction_list = contraction_list
            ctx.dense = dense
            ctx.save_for_backward(out, log_scale, *tensors)
            if dense:
                ctx.mark_non_differentiable(log_scale)
            return out, log_scale

        @staticmethod
        def backward(ctx, g_out, g_c):
            out, log_scale, *tensors = ctx.saved_tensors
            inputs, output = global_einsum(ctx.contraction_list, len(tensors))
            shapes = [tuple(t.shape) for t in tensors]
            with torch.no_grad():
                if ctx.dense:
                    # out = T itself: G = g_out, no scale to undo
                    g = torch.zeros_like(out) if g_out is None else g_out.to(out.dtype)
                    shift = None
                else:
                    g = split_output_cotangent(out, g_out, g_c)
                    shift = -log_scale.to(torch.float64)
                grads = []
                for k, t in enumerate(tensors):
                    if not ctx.needs_input_grad[3 + k]:
                        grads.append(None)
                        continue
                    einstr, spec = gradient_einsum(inputs, output, k, shapes)
                    ops = []
                    for item in spec:
                        if item[0] == "G":
                            ops.append(g)
                        elif item[0] == "X":
                            ops.append(tensors[item[1]].detach().to(g.dtype))
                        else:
                            ops.append(torch.ones(item[1], dtype=g.dtype, device=g.device))
                    if ops:
                        r_hat, c_r = contract_split(einstr, *ops)
                        c_tot = c_r.to(torch.float64) if shift is None else c_r.to(torch.float64) + shift
                    else:  # a single scalar-valued operand: dT/dX = 1
                        r_hat, c_tot = torch.ones_like(t, dtype=g.dtype), torch.zeros((), dtype=torch.float64, device=g.device)
                        if shift is not None:
                            c_tot = c_tot + shift
                    grad = r_hat * torch.exp(c_tot).to(r_hat.dtype)
                    if not output:  # scalar result: its cotangent is a number, not an operand
                        grad = grad * g.reshape(())
                    grads.append(grad.to(t.dtype).reshape(t.shape))
            return (None, None, None, *grads)

    return ContractFn
