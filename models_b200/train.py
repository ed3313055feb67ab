"""Training step of the DLRM path: forward with saved activations, binary cross-entropy, backward and
optimizer update as hand-written CUDA (include/mm_b200.h K14), behind the reference's `compile` / `fit` / `train_step`
(merlin/models/tf/models/base.py:1121-1231; optimizers: tf.keras.optimizers.{SGD, Adagrad, Adam}, LazyAdam
blocks/optimizer.py:342).

What a step launches (DLRM, bottom [.., D], top [...], BinaryOutput):

    forward   concat+split -> dense_tc per bottom layer (fp32 activation saved + operand of the next layer) -> fused
              lookup + interaction -> dense_tc per top layer.  D = 64 with table mirrors: the kernel reads operand-format
              rows and writes the top tower's split operand directly; otherwise fp32 rows + one split pass
    loss      mm_bce_head_fwd_bwd: output Dense(1) + sigmoid + BCE forward AND backward in one pass
    backward  per Dense layer mm_dense_wgrad[_split] (dW, db) + mm_dense_dgrad (input gradient, relu mask fused);
              mm_dlrm_interact_backward: pair gradients -> IndexedSlices per table + bottom-vector gradient
    update    mm_opt_tick, [DP: all-reduce of the dense gradient arena, all-gather of the slices], mm_dense_apply over the flat
              parameter arena, mm_sparse_rows_apply (duplicate ids summed, one update per touched row), mm_split_weights
              refresh of the tensor-core operand copies (in place)

All Dense variables of the model are re-homed into ONE flat fp32 arena (gradients and optimizer slots mirror its layout), so
the dense update is one launch and data-parallel training needs one all-reduce.  Every buffer is static: a step can be
captured into a CUDA graph (`DLRMTrainer.capture()` / `replay()`), the learning rate lives in device memory.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi, ops
from .blocks import DLRM, MLP, _Dense
from .core import default_device, get_feature

INT32_MAX = 2**31 - 1
DENSE_PATH_MAX_ROWS = 131072  # tables up to this size accumulate duplicate ids in a dense (rows, D) gradient


class History:
    """What Keras `fit` returns: `.history["loss"]` = mean batch loss of every epoch."""

    def __init__(self, history: Dict[str, List[float]]):
        self.history = history
        self.epoch = list(range(len(history.get("loss", []))))


class Optimizer:
    """Hyper-parameters of one of the update rules in include/mm_b200.h (Keras argument names and defaults)."""

    kind = "sgd"

    def __init__(self, learning_rate: float, beta_1: float = 0.0, beta_2: float = 0.0, epsilon: float = 1e-7,
                 initial_accumulator_value: float = 0.0):
        if learning_rate < 0:
            raise ValueError("learning_rate must be >= 0")
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.initial_accumulator_value = float(initial_accumulator_value)

    def hyper(self) -> np.ndarray:
        h = np.zeros(_cabi.HYPER_COUNT, dtype=np.float32)
        h[_cabi.HYPER_LR], h[_cabi.HYPER_BETA1], h[_cabi.HYPER_BETA2] = self.learning_rate, self.beta_1, self.beta_2
        h[_cabi.HYPER_EPS] = self.epsilon
        return h

    @property
    def slots(self) -> int:
        return {"sgd": 0, "adagrad": 1, "adam": 2}[self.kind]

    def get_config(self) -> dict:
        return {"name": type(self).__name__, "learning_rate": self.learning_rate, "beta_1": self.beta_1, "beta_2": self.beta_2,
                "epsilon": self.epsilon, "initial_accumulator_value": self.initial_accumulator_value}


class SGD(Optimizer):
    """tf.keras.optimizers.SGD(learning_rate=0.01) without momentum."""

    kind = "sgd"

    def __init__(self, learning_rate: float = 0.01, momentum: float = 0.0, **kwargs):
        if momentum:
            raise NotImplementedError("SGD momentum is not implemented")
        super().__init__(learning_rate)


class Adagrad(Optimizer):
    """tf.keras.optimizers.Adagrad(learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-7)."""

    kind = "adagrad"

    def __init__(self, learning_rate: float = 0.001, initial_accumulator_value: float = 0.1, epsilon: float = 1e-7, **kwargs):
        if initial_accumulator_value < 0:
            raise ValueError("initial_accumulator_value must be non-negative")
        super().__init__(learning_rate, epsilon=epsilon, initial_accumulator_value=initial_accumulator_value)


class Adam(Optimizer):
    """tf.keras.optimizers.Adam(0.001, 0.9, 0.999, 1e-7).  Embedding rows are updated lazily (only the rows a batch looked
    up, from the summed duplicate gradients) — the reference's LazyAdam (blocks/optimizer.py:342)."""

    kind = "adam"

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-7, **kwargs):
        super().__init__(learning_rate, beta_1, beta_2, epsilon)


LazyAdam = Adam

_BY_NAME = {"sgd": SGD, "adagrad": Adagrad, "adam": Adam, "lazyadam": Adam, "lazy_adam": Adam}


def get_optimizer(spec) -> Optimizer:
    if isinstance(spec, Optimizer):
        return spec
    if isinstance(spec, str):
        if spec.lower() not in _BY_NAME:
            raise ValueError(f"Unknown optimizer {spec!r}; supported: {sorted(_BY_NAME)}")
        return _BY_NAME[spec.lower()]()
    raise TypeError(f"optimizer must be a name or a models_b200.train.Optimizer, got {type(spec).__name__}")


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class DenseArena:
    """All Dense variables of a chain of layers in one flat fp32 buffer (kernel then bias per layer, 256-byte aligned);
    `grad`, `state1`, `state2` share the layout.  The layers' `kernel` / `bias` become views of it."""

    def __init__(self, layers: Sequence[_Dense], optimizer: Optimizer, device):
        self.layers = list(layers)
        off = 0
        self.kernel_off, self.bias_off = [], []
        for l in self.layers:
            if l.kernel is None:
                raise RuntimeError(f"{l.name}: build the model before compiling it for training")
            self.kernel_off.append(off)
            off = _align(off + l.kernel.numel())
            self.bias_off.append(off if l.bias is not None else None)
            if l.bias is not None:
                off = _align(off + l.bias.numel())
        self.size = off
        self.w = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.w)
        self.state1 = torch.full_like(self.w, optimizer.initial_accumulator_value) if optimizer.slots >= 1 else None
        self.state2 = torch.zeros_like(self.w) if optimizer.slots >= 2 else None
        for i, l in enumerate(self.layers):
            k = self.view(self.w, i, "kernel")
            k.copy_(l.kernel)
            l.kernel = k
            if l.bias is not None:
                b = self.view(self.w, i, "bias")
                b.copy_(l.bias)
                l.bias = b
            l._weights_changed()

    def view(self, buf: torch.Tensor, i: int, what: str) -> Optional[torch.Tensor]:
        l = self.layers[i]
        if what == "kernel":
            K, N = l.kernel.shape
            return buf[self.kernel_off[i]: self.kernel_off[i] + K * N].view(K, N)
        if self.bias_off[i] is None:
            return None
        return buf[self.bias_off[i]: self.bias_off[i] + l.units]


def gather_slices(ids: torch.Tensor, slices: torch.Tensor, group) -> tuple:
    """Data-parallel embedding gradients: Horovod all-gathers IndexedSlices (values and indices concatenated over the ranks,
    values divided by the world size; models/base.py:476-508).  ids (T, B) int32, slices (T, B, D) ->
    (T, world*B) ids and (T, world*B, D) slices, identical on every rank."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    T, B = ids.shape
    all_ids = torch.empty((world, T, B), dtype=ids.dtype, device=ids.device)
    all_sl = torch.empty((world,) + tuple(slices.shape), dtype=slices.dtype, device=slices.device)
    dist.all_gather([all_ids[r] for r in range(world)], ids.contiguous(), group=group)
    dist.all_gather([all_sl[r] for r in range(world)], slices.contiguous(), group=group)
    return (all_ids.permute(1, 0, 2).reshape(T, world * B).contiguous(),
            all_sl.permute(1, 0, 2, 3).reshape(T, world * B, slices.shape[2]).contiguous())


class DLRMTrainer:
    """Static-buffer training step of a DLRM RankingModel at one batch size."""

    def __init__(self, model, optimizer: Optimizer, batch_size: int, device=None, group=None):
        from .models import BinaryOutput

        body = model.body
        if not isinstance(body, DLRM) or body.top_block is None or body.bottom_block is None:
            raise NotImplementedError("train_step is implemented for DLRMModel with bottom and top blocks")
        if not isinstance(model.prediction, BinaryOutput):
            raise NotImplementedError("train_step needs a BinaryOutput head (binary cross-entropy)")
        if body.sharded is not None:
            raise NotImplementedError("training with row-sharded tables is not implemented (forward only)")
        if not body.can_emit_split():
            raise NotImplementedError("training needs <= 32 interaction features and embedding_dim in {16, 32, 64, 128}")
        self.model, self.body, self.opt = model, body, optimizer
        self.device = torch.device(device) if device is not None else default_device()
        self.B = int(batch_size)
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(group)
        if not model.built:
            model.build(self.device)
        for blk in (body.bottom_block, body.top_block):
            if not isinstance(blk, MLP) or blk.has_normalization or blk.dropout:
                raise NotImplementedError("training supports MLPBlock towers without normalization / dropout")
        self.bottom = body.bottom_block.dense_layers
        self.top = body.top_block.dense_layers
        self.head = model.prediction.to_call
        for l in self.bottom + self.top:
            if l.activation not in ("relu", "linear"):
                raise NotImplementedError(f"{l.name}: training supports relu / linear tower activations, got {l.activation!r}")
        if self.head.input_dim > 256:
            raise NotImplementedError("the output layer's input must be <= 256 wide")
        self.arena = DenseArena(self.bottom + self.top + [self.head], optimizer, self.device)
        nb, nt = len(self.bottom), len(self.top)
        self._wsplit = [ops.split_weights(l.kernel) for l in self.bottom + self.top]
        for l, ws in zip(self.bottom + self.top, self._wsplit):
            l._w_split = ws  # the forward path of this model keeps reading the refreshed operand copies
        self.hyper = torch.from_numpy(optimizer.hyper()).to(self.device)
        # layers wider than mm_dense_dgrad's 128 units take dX = dZ W^T through the tensor-core forward GEMM on the transposed
        # kernel: per such layer a transposed copy, its split operand and the split of dZ
        self._wide: Dict[int, dict] = {}
        for li, l in enumerate(self.bottom + self.top):
            if l.units > 128 and li != 0:  # the first bottom layer needs no input gradient
                K, N = l.kernel.shape
                wT = torch.empty((N, K), dtype=torch.float32, device=self.device)
                self._wide[li] = dict(wT=wT, wT_split=torch.zeros((ops.tc_padded_n(K), 2 * ops.tc_padded_k(N)), dtype=torch.bfloat16, device=self.device),
                                      dz_split=torch.zeros((self.B, 2 * ops.tc_padded_k(N)), dtype=torch.bfloat16, device=self.device))

        # ---- tables
        emb = body.embeddings
        self.feats = list(emb.feature_names)
        self.slots = body.slots()
        self.D = body.embedding_dim
        self.tables = [emb.feature_to_table[f] for f in self.feats]
        seen = set()
        for t in self.tables:
            if id(t) in seen:
                raise NotImplementedError("training with a table shared between features is not implemented")
            seen.add(id(t))
            if not t.trainable:
                raise NotImplementedError("frozen embedding tables are not implemented in the training step")
        T, B, D = len(self.tables), self.B, self.D
        self.slices = torch.zeros((T, B, D), dtype=torch.float32, device=self.device)
        self.rep = [ops.fill_i32(torch.empty(t.table.shape[0], dtype=torch.int32, device=self.device), INT32_MAX) for t in self.tables]
        self.tstate1 = [torch.full_like(t.table, optimizer.initial_accumulator_value) if optimizer.slots >= 1 else None for t in self.tables]
        self.tstate2 = [torch.zeros_like(t.table) if optimizer.slots >= 2 else None for t in self.tables]
        # tables with few rows (every id repeats many times per batch) sum their slices into a dense accumulator
        self.tdense = [torch.zeros_like(t.table) if t.table.shape[0] <= DENSE_PATH_MAX_ROWS else None for t in self.tables]

        # ---- activations and gradients
        # operand-format rows for the lookup + interaction kernels (forward and backward) (D = 64, mirrors enabled): the tables' mirrors (the
        # optimizer kernels keep them in step) and the bottom vector as the last bottom layer's split output
        from .blocks import table_mirror

        self.operand_rows = bool(D == 64 and table_mirror())
        if self.operand_rows:
            for t in self.tables:
                t.operand_mirror()
        f32 = dict(dtype=torch.float32, device=self.device)
        self.K0 = len(body.continuous.features)
        self.x0_split = torch.zeros((B, 2 * ops.tc_padded_k(self.K0)), dtype=torch.bfloat16, device=self.device)
        self.h = [torch.zeros((B, l.units), **f32) for l in self.bottom]
        n_split = len(self.bottom) if self.operand_rows else len(self.bottom) - 1
        self.h_split = [torch.zeros((B, 2 * ops.tc_padded_k(l.units)), dtype=torch.bfloat16, device=self.device) for l in self.bottom[:n_split]]
        F = len(self.slots)
        self.OW = D + F * (F - 1) // 2
        self.ldA = (self.OW + 3) // 4 * 4
        self.A = None if self.operand_rows else torch.zeros((B, self.ldA), **f32)
        self.dA = torch.zeros((B, self.ldA), **f32)
        self.A_split = torch.zeros((B, 2 * ops.tc_padded_k(self.OW)), dtype=torch.bfloat16, device=self.device)
        self.t = [torch.zeros((B, l.units), **f32) for l in self.top]
        self.t_split = [torch.zeros((B, 2 * ops.tc_padded_k(l.units)), dtype=torch.bfloat16, device=self.device) for l in self.top[:-1]]
        self.dt = [torch.zeros((B, l.units), **f32) for l in self.top]
        self.dh = [torch.zeros((B, l.units), **f32) for l in self.bottom]
        self.logits = torch.zeros(B, **f32)
        self.loss = torch.zeros(1, **f32)
        self.oob = emb.counter(self.device)
        self.steps = 0
        self._graph = None
        self._static: Optional[Dict[str, torch.Tensor]] = None
        self._static_y: Optional[torch.Tensor] = None

    # ---- one step on device tensors ------------------------------------------------------------------------------
    def _indices(self, inputs) -> List[torch.Tensor]:
        from .inputs import _as_index

        out = []
        for f in self.feats:
            i = get_feature(inputs, f)
            out.append(i if i.dtype in (torch.uint8, torch.uint16) else _as_index(i).reshape(-1))
        return out

    def forward_backward(self, inputs: Dict[str, torch.Tensor], targets: torch.Tensor, sample_weight=None) -> None:
        """Forward (activations saved), loss and backward: fills the gradient arena and the IndexedSlices.  Batches smaller
        than the compiled size run in the leading rows of the same buffers."""
        a = self.arena
        nb, nt = len(self.bottom), len(self.top)
        D = self.D
        self.loss.zero_()
        cont = self.body.continuous(inputs)
        pieces = [cont[k] for k in sorted(cont)]
        b = int(pieces[0].shape[0])
        if b > self.B or b < 1:
            raise ValueError(f"this trainer was compiled for batches of up to {self.B} samples, got {b}")
        if targets.numel() != b:
            raise ValueError(f"targets must hold {b} values, got {tuple(targets.shape)}")

        def v(t):
            return t[:b]

        h, t_, dt, dh = [v(x) for x in self.h], [v(x) for x in self.t], [v(x) for x in self.dt], [v(x) for x in self.dh]
        ops.concat_split(pieces, out=v(self.x0_split))  # the concatenated continuous columns exist only as this operand
        # -- bottom tower
        op, K = v(self.x0_split), self.K0
        for i, l in enumerate(self.bottom):
            nxt = v(self.h_split[i]) if i < len(self.h_split) else None
            ops.dense_tc(op, K, self._wsplit[i], l.units, l.bias, l.activation, out_f32=h[i], out_split=nxt)
            op, K = nxt, l.units
        # -- lookup + interaction (fp32 rows)
        idx = self._indices(inputs)
        tabs = [t.table for t in self.tables]
        rows = [t.shape[0] for t in tabs]
        tslots = [self.slots[f] for f in self.feats]
        bslot = self.slots["bottom_block"]
        dA_view = self.dA[:b, :self.OW]
        if self.operand_rows:
            # operand-format rows in, split-bf16 operand of the top tower out: no fp32 copy of [bottom | interactions] exists
            A_view = None
            ops.dlrm_lookup_interact([t._mirror for t in self.tables], idx, tslots, rows, D, v(self.h_split[-1]), bslot,
                                     v(self.A_split), self.oob, operand_rows=True)
        else:
            A_view = self.A[:b, :self.OW]
            ops.dlrm_lookup_interact(tabs, idx, tslots, rows, D, h[-1], bslot, A_view, self.oob)
            ops.split_rows(A_view, out=v(self.A_split))
        # -- top tower
        op, K = v(self.A_split), self.OW
        for i, l in enumerate(self.top):
            nxt = v(self.t_split[i]) if i < nt - 1 else None
            ops.dense_tc(op, K, self._wsplit[nb + i], l.units, l.bias, l.activation, out_f32=t_[i], out_split=nxt)
            op, K = nxt, l.units
        # -- output layer + loss, forward and backward
        hi = nb + nt
        ops.bce_head_fwd_bwd(t_[-1], self.head.kernel.reshape(-1), self.head.bias, targets.reshape(-1), self.loss, dt[-1],
                             a.view(a.grad, hi, "kernel").reshape(-1), a.view(a.grad, hi, "bias"),
                             mask_relu=self.top[-1].activation == "relu", sample_weight=sample_weight, logits=v(self.logits))
        # -- top tower backward
        for i in range(nt - 1, -1, -1):
            l = self.top[i]
            if i == 0 and A_view is None:
                ops.dense_wgrad_split(v(self.A_split), self.OW, dt[0], a.view(a.grad, nb, "kernel"), a.view(a.grad, nb, "bias"))
            else:
                ops.dense_wgrad(t_[i - 1] if i > 0 else A_view, dt[i], a.view(a.grad, nb + i, "kernel"), a.view(a.grad, nb + i, "bias"))
            if i > 0:
                self._dgrad(nb + i, l, dt[i], dt[i - 1], t_[i - 1] if self.top[i - 1].activation == "relu" else None)
            else:
                self._dgrad(nb, l, dt[0], dA_view, None)
        # -- interaction + lookup backward
        self._slices = [self.slices[t][:b] for t in range(len(tabs))]
        if self.operand_rows:
            ops.dlrm_interact_backward([t._mirror for t in self.tables], idx, tslots, rows, D, v(self.h_split[-1]), bslot, dA_view,
                                       self._slices, dh[-1], mask_bottom=self.bottom[-1].activation == "relu", operand_rows=True)
        else:
            ops.dlrm_interact_backward(tabs, idx, tslots, rows, D, h[-1], bslot, dA_view, self._slices, dh[-1],
                                       mask_bottom=self.bottom[-1].activation == "relu")
        # -- bottom tower backward
        for i in range(nb - 1, -1, -1):
            l = self.bottom[i]
            if i > 0:
                ops.dense_wgrad(h[i - 1], dh[i], a.view(a.grad, i, "kernel"), a.view(a.grad, i, "bias"))
            else:
                ops.dense_wgrad_split(v(self.x0_split), self.K0, dh[0], a.view(a.grad, 0, "kernel"), a.view(a.grad, 0, "bias"))
            if i > 0:
                self._dgrad(i, l, dh[i], dh[i - 1], h[i - 1] if self.bottom[i - 1].activation == "relu" else None)
        self._idx, self._b = idx, b

    def _dgrad(self, li: int, layer: _Dense, dz: torch.Tensor, dx: torch.Tensor, mask: Optional[torch.Tensor]) -> None:
        """dx = dz W^T (zeroed where mask <= 0) for layer `li` of bottom + top."""
        wide = self._wide.get(li)
        if wide is None:
            ops.dense_dgrad(dz, layer.kernel, dx, mask=mask)
            return
        K, N = layer.kernel.shape
        b = dz.shape[0]
        wide["wT"].copy_(layer.kernel.t())
        ops.split_weights(wide["wT"], out=wide["wT_split"])
        ops.split_rows(dz, out=wide["dz_split"][:b])
        ops.dense_tc(wide["dz_split"][:b], N, wide["wT_split"], K, None, None, out_f32=dx)
        if mask is not None:
            ops.relu_mask(dx, mask)

    def apply_gradients(self) -> None:
        a = self.arena
        ops.opt_tick(self.hyper)
        idx, slices, Bt = self._idx, self._slices, self._b
        scale = 1.0
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(a.grad, group=self.group)
            scale = 1.0 / self.world
            ids32 = torch.stack([ops.widen_index(i).to(torch.int32) for i in idx])
            all_ids, all_sl = gather_slices(ids32, torch.stack(slices), self.group)
            all_sl.mul_(scale)
            idx = [all_ids[t] for t in range(all_ids.shape[0])]
            slices = [all_sl[t] for t in range(all_sl.shape[0])]
            Bt = self._b * self.world
        ops.dense_apply(self.opt.kind, a.w, a.grad, a.state1, a.state2, self.hyper, grad_scale=scale)
        tabs = []
        for t, tb in enumerate(self.tables):
            mirror = tb._mirror if (tb._mirror is not None and tb._mirror.shape[0] == tb.table.shape[0]) else None
            tabs.append(dict(weights=tb.table, indices=idx[t], grad_rows=slices[t], rep_map=self.rep[t], state1=self.tstate1[t],
                             state2=self.tstate2[t], mirror=mirror, dense_grad=self.tdense[t]))
        ops.sparse_rows_apply(self.opt.kind, tabs, Bt, self.D, self.hyper)
        for l, ws in zip(self.bottom + self.top, self._wsplit):
            ops.split_weights(l.kernel, out=ws)

    def step(self, inputs: Dict[str, torch.Tensor], targets: torch.Tensor, sample_weight=None) -> torch.Tensor:
        """One eager training step; returns the batch loss as a (1,) device tensor (valid until the next step)."""
        self.forward_backward(inputs, targets, sample_weight)
        self.apply_gradients()
        self._after_step()
        return self.loss

    def _after_step(self) -> None:
        from .core import bump_weights_version

        self.steps += 1
        self.head._bias_host = None
        bump_weights_version()  # forward graphs captured earlier hold scalars / operand copies of the old variables

    # ---- CUDA-graph replay over static input buffers ------------------------------------------------------------
    def capture(self, inputs: Dict[str, torch.Tensor], targets: torch.Tensor, clone: bool = True) -> None:
        """Capture forward + backward + update into ONE CUDA graph over copies of `inputs` / `targets` (single GPU; with a
        process group the collectives stay eager between two graphs).  clone=False: the given tensors ARE the static
        buffers (e.g. views of one packed device buffer that a single H2D copy refreshes before replay())."""
        if self.world > 1:
            raise NotImplementedError("graph capture of the data-parallel step is not implemented")
        self._static = {k: (v.clone() if clone else v) for k, v in inputs.items()}
        self._static_y = targets.clone() if clone else targets
        self.model.defer_index_check(True)
        try:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            # warm-up steps DO train: snapshot and restore every variable so that capture has no side effect
            snap = self._snapshot()
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.forward_backward(self._static, self._static_y)
                    self.apply_gradients()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(self._graph):
                self.forward_backward(self._static, self._static_y)
                self.apply_gradients()
            self.launches_per_step = ops.launch_count() - n0
            self._restore(snap)
        finally:
            self.model.defer_index_check(False)

    def _snapshot(self):
        return dict(w=self.arena.w.clone(), s1=None if self.arena.state1 is None else self.arena.state1.clone(),
                    s2=None if self.arena.state2 is None else self.arena.state2.clone(), hyper=self.hyper.clone(),
                    tables=[t.table.clone() for t in self.tables],
                    ts1=[None if s is None else s.clone() for s in self.tstate1], ts2=[None if s is None else s.clone() for s in self.tstate2])

    def _restore(self, snap) -> None:
        self.arena.w.copy_(snap["w"])
        self.arena.grad.zero_()
        if snap["s1"] is not None:
            self.arena.state1.copy_(snap["s1"])
        if snap["s2"] is not None:
            self.arena.state2.copy_(snap["s2"])
        self.hyper.copy_(snap["hyper"])
        for t, w, s1, s2, a1, a2 in zip(self.tables, snap["tables"], snap["ts1"], snap["ts2"], self.tstate1, self.tstate2):
            t.table.copy_(w)
            if s1 is not None:
                a1.copy_(s1)
            if s2 is not None:
                a2.copy_(s2)
            if t._mirror is not None and t._mirror.shape[0] == t.table.shape[0]:
                ops.split_rows(t.table, out=t._mirror)
        for l, ws in zip(self.bottom + self.top, self._wsplit):
            ops.split_weights(l.kernel, out=ws)

    def replay(self, inputs: Optional[Dict[str, torch.Tensor]] = None, targets: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self._graph is None:
            raise RuntimeError("capture() first")
        if inputs is not None:
            for k, v in self._static.items():
                v.copy_(inputs[k], non_blocking=True)
        if targets is not None:
            self._static_y.copy_(targets.reshape(self._static_y.shape), non_blocking=True)
        self._graph.replay()
        self._after_step()
        return self.loss

    def check_indices(self) -> None:
        """Raise IndexError if any batch since the last check carried an id outside its table (the forward read a zero
        row for it and its gradient was dropped).  One device-to-host read: `fit` calls it once per epoch, not per step."""
        if self.oob is None:
            return
        n = int(self.oob.item())
        if n:
            self.oob.zero_()
            raise IndexError(f"{n} indices out of range for the embedding tables "
                             "(TF raises InvalidArgumentError: indices[...] is not in [0, rows))")

    def set_learning_rate(self, lr: float) -> None:
        self.opt.learning_rate = float(lr)
        self.hyper[_cabi.HYPER_LR] = float(lr)

    def gradients(self) -> Dict[str, torch.Tensor]:
        """Dense gradients by variable name (after forward_backward, before apply_gradients) — for parity tests."""
        out = {}
        for i, l in enumerate(self.arena.layers):
            out[f"{l.name}/kernel"] = self.arena.view(self.arena.grad, i, "kernel")
            b = self.arena.view(self.arena.grad, i, "bias")
            if b is not None:
                out[f"{l.name}/bias"] = b
        return out
