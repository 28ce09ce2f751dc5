"""Packed (bf16, kernel-layout) copies of the fp32 master parameters, rebuilt lazily when a parameter's version changes
(optimizer.step / load_state_dict bump ``Tensor._version``), never stored in the state_dict."""
import torch

from . import ops


def _ver(*ts):
    return tuple((t.data_ptr(), t._version) for t in ts if t is not None)


class PackedConv:
    """conv (+ FrozenBN) -> wf bf16 [Cout, ldk] with BN scale folded, wd bf16 [Cin, taps*Cout] (flipped taps) for dgrad,
    scale / bias fp32 [Cout]."""

    def __init__(self, conv, bn=None, need_dgrad=True, ldk=None):
        self.conv, self.bn = conv, bn
        w = conv.weight
        self.Cout, self.Cin, self.kh, self.kw = w.shape
        self.taps = self.kh * self.kw
        self.K = self.taps * self.Cin
        self.ldk = ldk or self.K
        self.need_dgrad = need_dgrad
        self.trainable = w.requires_grad
        self._key = None
        self.wf = self.wd = self.scale = self.bias = None

    def current_key(self):
        bn = (self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var) if self.bn is not None else ()
        return _ver(self.conv.weight, getattr(self.conv, "bias", None), *bn)

    def refresh(self):
        w = self.conv.weight
        bn = None
        if self.bn is not None:
            bn = (self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        cb = getattr(self.conv, "bias", None)
        key = _ver(w, cb, *(bn or ()))
        if key == self._key:
            return
        dev = w.device
        if self.wf is None or self.wf.device != dev:
            self.wf = torch.empty(self.Cout, self.ldk, dtype=ops.t16(), device=dev)
            self.wd = torch.empty(self.Cin, self.taps * self.Cout, dtype=ops.t16(), device=dev) if self.need_dgrad else None
            self.scale = torch.empty(self.Cout, dtype=torch.float32, device=dev)
            self.bias = torch.empty(self.Cout, dtype=torch.float32, device=dev)
        ops.pack_conv(w.detach(), bn, cb.detach() if cb is not None else None, self.wf, self.ldk, self.wd, self.scale, self.bias,
                      eps=self.bn.eps if self.bn is not None else 1e-5)
        self._key = key


class PackedStem(PackedConv):
    """The 7x7 stem convolution: ``wf`` as for every convolution plus ``wrow`` [7 tap rows][4 K chunks][64][8], the layout
    rb_stem_pool reads (K index 4 * s + c over 8 pixels x 4 channels of a 16-bit HWC4 image row; pixel 7 / channel 3 are zero)."""

    wrow = None

    def refresh(self):
        key = self._key
        super().refresh()
        if self._key == key and self.wrow is not None:
            return
        dev = self.wf.device
        k = torch.arange(32, device=dev)
        s_, c_ = k // 4, k % 4
        valid = (s_ < 7) & (c_ < 3)
        r = torch.arange(7, device=dev).view(7, 1)
        col = ((r * 7 + s_.clamp(max=6)) * 3 + c_.clamp(max=2)).reshape(-1)         # [7 * 32] columns of wf, order (r, s, c)
        w = self.wf[:, col].view(self.Cout, 7, 32) * valid.view(1, 1, 32).to(self.wf.dtype)
        self.wrow = w.permute(1, 2, 0).reshape(7, 4, 8, self.Cout).permute(0, 1, 3, 2).contiguous()


class PackedLinear:
    """weight fp32 [N,K] -> wb bf16 [Npad,K], wt bf16 [K,Npad]; bias fp32 [Npad].  ``pad_to`` zero-pads the output dim
    (used for the 4-wide box head so every pitch is a multiple of 16 bytes)."""

    def __init__(self, weight, bias=None, pad_to=None, hilo=False):
        """``hilo``: also keep ``wb2`` [Npad, 2K] = (weight rounded to 16 bits | rounding residual) for a two-tap forward GEMM that
        multiplies by the weight to ~2^-22 (``lin_taps``)."""
        self.weight, self.bias_p = weight, bias
        self.N, self.K = weight.shape
        self.Np = max(self.N, pad_to or 0)
        self._key = None
        self.hilo = hilo
        self.wb = self.wt = self.bias = self.wb2 = None

    def current_key(self):
        return _ver(self.weight, self.bias_p)

    def refresh(self):
        key = _ver(self.weight, self.bias_p)
        if key == self._key:
            return
        dev = self.weight.device
        if self.wb is None or self.wb.device != dev:
            self.wb = torch.zeros(self.Np, self.K, dtype=ops.t16(), device=dev)
            self.wt = torch.zeros(self.K, self.Np, dtype=ops.t16(), device=dev)
            if self.Np != self.N:
                self.bias = torch.zeros(self.Np, dtype=torch.float32, device=dev)
        ops.pack_linear(self.weight.detach(), self.wb, self.wt)
        if self.hilo:
            if self.wb2 is None or self.wb2.device != dev:
                self.wb2 = torch.zeros(self.Np, 2 * self.K, dtype=ops.t16(), device=dev)
            ops.pack_linear_hilo(self.weight.detach(), self.wb2)
        if self.bias_p is not None:
            if self.Np != self.N:
                self.bias[:self.N].copy_(self.bias_p.detach())
            else:
                self.bias = self.bias_p.detach()
        self._key = key


def lin_taps(pack, rows=None):
    """(B operand, taps) of the FORWARD GEMM of a packed linear layer: the plain 16-bit weight with one tap, or -- for ``hilo`` packs
    -- the (hi | residual) pair with two taps that read the same A rows.  ``rows``: slice of the output features."""
    w = pack.wb2 if getattr(pack, "wb2", None) is not None else pack.wb
    if rows is not None:
        w = w[rows]
    taps = ((0, 0), (0, pack.K)) if getattr(pack, "wb2", None) is not None else ((0, 0),)
    return w, taps


class PackedStack:
    """Row-blocks of several weights packed side by side: wb [n*Nblk, K] and wt [K, n*Nblk] (decoder cross-attention K / V
    projections of all layers, so the memory is projected by ONE GEMM)."""

    def __init__(self, weights, biases, row0, nrows, hilo=False):
        self.weights, self.biases, self.row0, self.nrows = weights, biases, row0, nrows
        self.K = weights[0].shape[1]
        self.n = len(weights)
        self._key = None
        self.hilo = hilo
        self.wb = self.wt = self.bias = self.wb2 = None

    def current_key(self):
        return _ver(*self.weights, *self.biases)

    def refresh(self):
        key = _ver(*self.weights, *self.biases)
        if key == self._key:
            return
        dev = self.weights[0].device
        N = self.n * self.nrows
        if self.wb is None or self.wb.device != dev:
            self.wb = torch.empty(N, self.K, dtype=ops.t16(), device=dev)
            self.wt = torch.empty(self.K, N, dtype=ops.t16(), device=dev)
            self.bias = torch.empty(N, dtype=torch.float32, device=dev)
        if self.hilo and (self.wb2 is None or self.wb2.device != dev):
            self.wb2 = torch.empty(N, 2 * self.K, dtype=ops.t16(), device=dev)
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            sl = slice(i * self.nrows, (i + 1) * self.nrows)
            ops.pack_linear(w.detach()[self.row0:self.row0 + self.nrows], self.wb[sl], self.wt[:, sl])
            if self.hilo:
                ops.pack_linear_hilo(w.detach()[self.row0:self.row0 + self.nrows], self.wb2[sl])
            self.bias[sl].copy_(b.detach()[self.row0:self.row0 + self.nrows])
        self._key = key
