"""Helpers shared by the GPU parity tests (oracle = checker only)."""
import numpy as np


def unpack_golden(g, prefix):
    Ws, bs, i = [], [], 0
    while f'{prefix}_W{i}' in g:
        Ws.append(g[f'{prefix}_W{i}']); bs.append(g[f'{prefix}_b{i}']); i += 1
    return Ws, bs


def layers_of(Ws):
    return [Ws[0].shape[0]] + [w.shape[1] for w in Ws]


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def per_layer_grad_err(g_cuda, g_ref, layers):
    """max-abs error of each W_l / b_l block relative to that block's max-abs reference value."""
    out, o = [], 0
    for l in range(len(layers) - 1):
        n = layers[l] * layers[l + 1]
        out.append(('W%d' % l, rel_err(g_cuda[o:o + n], g_ref[o:o + n]))); o += n
    for l in range(len(layers) - 1):
        n = layers[l + 1]
        out.append(('b%d' % l, rel_err(g_cuda[o:o + n], g_ref[o:o + n]))); o += n
    return out


def random_biases(bs, seed, scale=0.1):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(b.shape) * scale for b in bs]
