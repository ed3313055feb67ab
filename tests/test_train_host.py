"""Host-side logic of the training step (models_b200/train.py) that needs no GPU: optimizer descriptors (Keras argument
names and defaults), the flat parameter arena, `compile` dispatch, the History object across the checkpoint pickle."""
import pickle

import numpy as np
import pytest
import torch

import models_b200 as mm
from models_b200 import _cabi, datasets, train
from models_b200.blocks import _Dense


def test_optimizer_descriptors_follow_keras_defaults():
    assert mm.SGD().learning_rate == 0.01 and mm.SGD().slots == 0
    a = mm.Adagrad()
    assert (a.learning_rate, a.initial_accumulator_value, a.epsilon, a.slots) == (0.001, 0.1, 1e-7, 1)
    ad = mm.Adam()
    assert (ad.learning_rate, ad.beta_1, ad.beta_2, ad.epsilon, ad.slots) == (0.001, 0.9, 0.999, 1e-7, 2)
    assert mm.LazyAdam is mm.Adam  # embedding rows are always updated lazily here
    h = mm.Adam(0.02, beta_1=0.8).hyper()
    assert h.shape == (_cabi.HYPER_COUNT,) and h.dtype == np.float32
    assert h[_cabi.HYPER_LR] == np.float32(0.02) and h[_cabi.HYPER_BETA1] == np.float32(0.8) and h[_cabi.HYPER_STEP] == 0
    assert train.get_optimizer("AdaGrad").kind == "adagrad" and train.get_optimizer("lazyadam").kind == "adam"
    with pytest.raises(ValueError, match="Unknown optimizer"):
        train.get_optimizer("ftrl")
    with pytest.raises(TypeError):
        train.get_optimizer(3)
    with pytest.raises(NotImplementedError):
        mm.SGD(0.1, momentum=0.9)
    with pytest.raises(ValueError):
        mm.Adagrad(-1.0)
    cfg = mm.Adagrad(0.05).get_config()
    assert cfg["name"] == "Adagrad" and cfg["learning_rate"] == 0.05


def test_dense_arena_rehomes_variables_as_views():
    layers = [_Dense(8, activation="relu"), _Dense(3, activation="linear", use_bias=False), _Dense(1, activation="sigmoid")]
    width = 5
    for l in layers:
        l.kernel = torch.randn(width, l.units)
        l.bias = torch.randn(l.units) if l.use_bias else None
        l.input_dim, l.built = width, True
        width = l.units
    before = [(l.kernel.clone(), None if l.bias is None else l.bias.clone()) for l in layers]
    arena = train.DenseArena(layers, mm.Adagrad(0.1), torch.device("cpu"))
    assert arena.size % 64 == 0 and arena.state1 is not None and arena.state2 is None
    assert float(arena.state1.min()) == pytest.approx(0.1)
    for i, (l, (k, b)) in enumerate(zip(layers, before)):
        assert torch.equal(l.kernel, k) and l.kernel.is_contiguous()
        assert l.kernel.data_ptr() == arena.view(arena.w, i, "kernel").data_ptr()  # the layer's variable IS the arena
        assert (l.kernel.data_ptr() - arena.w.data_ptr()) % 256 == 0
        if b is None:
            assert arena.view(arena.grad, i, "bias") is None
        else:
            assert torch.equal(l.bias, b) and l.bias.data_ptr() == arena.view(arena.w, i, "bias").data_ptr()
    arena.w.zero_()  # an update of the arena is an update of every layer
    assert all(float(l.kernel.abs().max()) == 0.0 for l in layers)
    g = arena.view(arena.grad, 0, "kernel")
    g.fill_(2.0)
    assert float(arena.grad.sum()) == 2.0 * layers[0].kernel.numel()


def test_compile_dispatches_on_its_argument():
    schema = datasets.criteo_schema({k: min(v, 50) for k, v in datasets.CRITEO_MAX.items()})
    model = mm.DLRMModel(schema, embedding_dim=16, bottom_block=mm.MLPBlock([32, 16]), top_block=mm.MLPBlock([16]))
    assert model.compile(optimizer="adagrad") is None and model.optimizer.kind == "adagrad"
    model.compile("sgd")
    assert model.optimizer.kind == "sgd"
    model.compile(mm.Adam(0.5))
    assert model.optimizer.learning_rate == 0.5
    model.compile()  # Keras-style bare compile: the default optimizer
    assert model.optimizer.kind == "adam"
    with pytest.raises(ValueError, match="either an example batch"):
        model.compile({"C1": np.zeros(4, np.int32)}, optimizer="sgd")
    with pytest.raises(NotImplementedError, match="binary"):
        model.compile(optimizer="sgd", loss="categorical_crossentropy")
    model.optimizer = None
    with pytest.raises(RuntimeError, match="compile"):
        model.train_step(({}, np.zeros(1)))
    with pytest.raises(RuntimeError, match="compile"):
        model.trainer(8)


def test_history_survives_the_structure_pickle():
    h = train.History({"loss": [0.7, 0.6]})
    g = pickle.loads(pickle.dumps(h))
    assert g.history == {"loss": [0.7, 0.6]} and g.epoch == [0, 1]
