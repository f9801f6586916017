"""Test helper: a CPU engine with the dp_* interface of WakeWordMLPModel, backed by the oracle classifier (torch autograd)."""
import numpy as np
import torch

from heybuddy_b200 import spec
from oracle import classifier as ocls


class OracleEngine:
    def __init__(self, seed: int = 11):
        self.params = {k: np.asarray(v, dtype=np.float64) for k, v in spec.init_classifier_weights(seed=seed).items()}
        self.names = list(self.params)
        self.m = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.t = 0
        self.grads = None

    def dp_select(self, x, y, thr):
        self._xy = (x.numpy(), y.numpy())
        prob = ocls.forward(self._xy[0], self.params)
        p = prob.reshape(-1)
        sel = ((self._xy[1] == 0) & (p >= thr)) | ((self._xy[1] == 1) & (p < 1 - thr))
        stats = torch.zeros(4, dtype=torch.float32)
        stats[1] = float(sel.sum())
        return torch.from_numpy(prob.astype(np.float32)), stats

    def dp_backward(self, n_total, negative_weight, thr, min_selected):
        x, y = self._xy
        prob, loss, n_sel, grads = ocls.forward_backward_torch(x, y, self.params, negative_weight, thr)
        scale = n_sel / float(n_total[0]) if float(n_total[0]) > 0 else 0.0     # mean over local rows -> share of the global mean
        self.grads = {k: g * scale for k, g in grads.items()}
        stats = torch.zeros(4, dtype=torch.float32)
        stats[0] = loss * scale
        stats[1] = n_sel
        stats[2] = 1.0 if float(n_total[0]) >= min_selected else 0.0
        stats[3] = n_sel / x.shape[0]
        return stats

    def dp_grads(self, buf=None, to_model=False):
        if to_model:
            flat = buf.numpy().astype(np.float64)
            o = 0
            for k in self.names:
                n = self.grads[k].size
                self.grads[k] = flat[o:o + n].reshape(self.grads[k].shape)
                o += n
            return buf
        return torch.from_numpy(np.concatenate([self.grads[k].reshape(-1) for k in self.names]).astype(np.float32))

    # the one-collective form: unnormalised sums [gradients | loss sum | rows selected], divided after the exchange
    def dp_local_step(self, x, y, negative_weight, thr):
        prob, stats = self.dp_select(x, y, thr)
        xn, yn = self._xy
        _, loss, n_sel, grads = ocls.forward_backward_torch(xn, yn, self.params, negative_weight, thr)
        self.grads = grads
        flat = np.concatenate([grads[k].reshape(-1) * n_sel for k in self.names] + [np.array([loss * n_sel, n_sel], dtype=np.float64)])
        stats[3] = n_sel / xn.shape[0]
        return prob, stats, torch.from_numpy(flat.astype(np.float32))

    def dp_apply(self, exchange, lr, min_selected, stats):
        flat = exchange.numpy().astype(np.float64)
        n_total = flat[-1]
        scale = 1.0 / n_total if n_total > 0 else 0.0
        o = 0
        for k in self.names:
            n = self.grads[k].size
            self.grads[k] = flat[o:o + n].reshape(self.grads[k].shape) * scale
            o += n
        stats[0] = flat[-2] * scale
        stats[1] = n_total
        stats[2] = 1.0 if n_total >= min_selected else 0.0
        self.dp_adam(lr, stats)
        return stats

    def dp_adam(self, lr, stats):
        if float(stats[2]) == 0.0:
            return
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        for k in self.names:
            g = self.grads[k]
            self.m[k] = b1 * self.m[k] + (1 - b1) * g
            self.v[k] = b2 * self.v[k] + (1 - b2) * g * g
            mh, vh = self.m[k] / (1 - b1 ** self.t), self.v[k] / (1 - b2 ** self.t)
            self.params[k] = self.params[k] - lr * mh / (np.sqrt(vh) + eps)

    def flat_params(self):
        return np.concatenate([self.params[k].reshape(-1) for k in self.names])
