"""Fast CPU variant of the oracle — TEST / BASELINE INFRASTRUCTURE ONLY (see keras_tf_oracle.py).

The same maths as `keras_tf_oracle.Wav2LetterOracle` (Keras Conv1D "same" tower, softmax,
log(p+1e-8) -> CTC with blank = V-1, batch-mean objective, Keras-2 Adam; reference
net.py:291-341,389,402-406,132) expressed with torch-CPU ops (mkldnn conv, native CTC,
autograd) so it can use every host core.  It is the "CPU restatement of the Keras/TF path
(Keras/TF unavailable offline)" that bench.py times as `cpu_baseline` / `--impl reference`
(BASELINE.md §3), and tests/test_oracle_crosscheck.py checks it against the numpy oracle.
"""
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle.keras_tf_oracle import EPSILON, same_padding, wav2letter_layer_specs


class TorchCpuWav2Letter:
    def __init__(self, input_size: int, grapheme_set_size: int, main_filter_count: int = 250,
                 out_filter_count: int = 2000, seed: int = 0, dtype=torch.float32, lr: float = 1e-4,
                 use_raw_wave_input: bool = False):
        self.specs = wav2letter_layer_specs(input_size, grapheme_set_size, main_filter_count, out_filter_count,
                                            use_raw_wave_input)
        self.dtype = dtype
        g = torch.Generator().manual_seed(seed)
        self.kernels: List[torch.Tensor] = []  # Keras layout (k, Cin, Cout)
        self.biases: List[torch.Tensor] = []
        for (_, cin, cout, k, _, _) in self.specs:
            limit = float(np.sqrt(6.0 / (k * cin + k * cout)))
            w = (torch.rand((k, cin, cout), generator=g, dtype=torch.float64) * 2 - 1) * limit
            self.kernels.append(w.to(dtype).requires_grad_(True))
            self.biases.append(torch.zeros(cout, dtype=dtype, requires_grad=True))
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, 0.9, 0.999, 1e-8
        self.iterations = 0
        self.m = [torch.zeros_like(p) for p in self.kernels + self.biases]
        self.v = [torch.zeros_like(p) for p in self.kernels + self.biases]

    def set_weights(self, kernels: Sequence[np.ndarray], biases: Sequence[np.ndarray]) -> None:
        self.kernels = [torch.as_tensor(np.asarray(w), dtype=self.dtype).clone().requires_grad_(True) for w in kernels]
        self.biases = [torch.as_tensor(np.asarray(b), dtype=self.dtype).clone().requires_grad_(True) for b in biases]

    def forward(self, x: torch.Tensor, return_logits: bool = False):
        a = x.to(self.dtype).transpose(1, 2)  # (B, C, T) for F.conv1d
        logits = None
        for (_, _, _, k, stride, act), w, b in zip(self.specs, self.kernels, self.biases):
            T = a.shape[2]
            _, pad_l, pad_r = same_padding(T, k, stride)
            a = F.conv1d(F.pad(a, (pad_l, pad_r)), w.permute(2, 1, 0), b, stride=stride)
            if act == "relu":
                a = torch.relu(a)
            else:
                logits = a.transpose(1, 2)
                a = torch.softmax(a, dim=1)
        probs = a.transpose(1, 2)  # (B, T', V)
        return (probs, logits) if return_logits else probs

    def losses(self, probs: torch.Tensor, labels: np.ndarray, prediction_lengths, label_lengths) -> torch.Tensor:
        """Per-utterance -log p(label | x), K.ctc_batch_cost semantics."""
        V = probs.shape[2]
        lp = torch.log_softmax(torch.log(probs + EPSILON), dim=2).transpose(0, 1)  # (T', B, V)
        ll = torch.as_tensor(np.asarray(label_lengths).reshape(-1), dtype=torch.long)
        pl = torch.as_tensor(np.asarray(prediction_lengths).reshape(-1), dtype=torch.long)
        targets = torch.as_tensor(np.maximum(np.asarray(labels), 0), dtype=torch.long)
        return F.ctc_loss(lp, targets, pl, ll, blank=V - 1, reduction="none", zero_infinity=False)

    def train_step(self, x: np.ndarray, labels: np.ndarray, prediction_lengths, label_lengths) -> float:
        """forward + CTC + backward + Keras-2 Adam; returns the batch-mean loss (net.py:389)."""
        params = self.kernels + self.biases
        for p in params:
            p.grad = None
        probs = self.forward(torch.as_tensor(x))
        loss = self.losses(probs, labels, prediction_lengths, label_lengths).mean()
        loss.backward()
        self.iterations += 1
        t = self.iterations
        lr_t = self.lr * np.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        with torch.no_grad():
            for i, p in enumerate(params):
                g = p.grad
                self.m[i].mul_(self.beta_1).add_(g, alpha=1.0 - self.beta_1)
                self.v[i].mul_(self.beta_2).addcmul_(g, g, value=1.0 - self.beta_2)
                p.sub_(lr_t * self.m[i] / (self.v[i].sqrt() + self.epsilon))
        return float(loss.item())

    def gradients(self, x: np.ndarray, labels: np.ndarray, prediction_lengths, label_lengths):
        params = self.kernels + self.biases
        for p in params:
            p.grad = None
        probs = self.forward(torch.as_tensor(x))
        per_example = self.losses(probs, labels, prediction_lengths, label_lengths)
        per_example.mean().backward()
        n = len(self.kernels)
        return (per_example.detach().numpy(), [p.grad.numpy() for p in params[:n]],
                [p.grad.numpy() for p in params[n:]])
