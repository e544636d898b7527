"""The GNNML3 model wrappers of the reference's four configured scripts, as one parameterised module with
the reference's sub-module names (``conv1..convL``, ``fc1``, ``fc2``), creation order and forward:

* ``graph8c``  -- graph8c.py:249-279      3 x ML3Layer(32||16), add-pool, tanh(fc1 -> 10)
* ``zinc``     -- Zinc12k.py:310-345      4 x ML3Layer(30||2),  add-pool, fc2(relu(fc1 -> 32)) -> 1
* ``exp``      -- exp_classify.py:264-295 3 x ML3Layer(32||16), mean-pool, fc2(relu(fc1 -> 10)) -> 1
* ``counting`` -- counting.py:335-372     5 x ML3Layer(16||16), add-pool, fc2(relu(fc1 -> 32)) -> 1

The same (unlearned) ``edge_attr2`` feeds every layer; the graph plan (CSR) is built once per batch.
"""
import torch
import torch.nn as nn

from .libs.spect_conv import ML3Layer, _LinearFn, _PRECISIONS
from .pool import global_add_pool, global_mean_pool

MODEL_CONFIGS = {
    "graph8c": dict(nlayer=3, nout1=32, nout2=16, head=(10,), pool="add", final="tanh"),
    "zinc": dict(nlayer=4, nout1=30, nout2=2, head=(32, 1), pool="add", final=None),
    "exp": dict(nlayer=3, nout1=32, nout2=16, head=(10, 1), pool="mean", final=None),
    "counting": dict(nlayer=5, nout1=16, nout2=16, head=(32, 1), pool="add", final=None),
}


class GNNML3(nn.Module):
    def __init__(self, config, ne, ninp, precision="fp32"):
        super().__init__()
        c = dict(MODEL_CONFIGS[config]) if isinstance(config, str) else dict(config)
        self.cfg = c
        self.precision = precision
        nin = c["nout1"] + c["nout2"]
        for l in range(c["nlayer"]):
            setattr(self, "conv%d" % (l + 1),
                    ML3Layer(learnedge=True, nedgeinput=ne, nedgeoutput=ne, ninp=ninp if l == 0 else nin,
                             nout1=c["nout1"], nout2=c["nout2"], precision=precision))
        self.fc1 = nn.Linear(nin, c["head"][0])
        if len(c["head"]) > 1:
            self.fc2 = nn.Linear(c["head"][0], c["head"][1])

    def forward(self, data):
        x = data.x
        edge_index = data.edge_index2
        edge_attr = data.edge_attr2
        for l in range(self.cfg["nlayer"]):
            x = getattr(self, "conv%d" % (l + 1))(x, edge_index, edge_attr)
        pool = global_add_pool if self.cfg["pool"] == "add" else global_mean_pool
        x = pool(x, data.batch, getattr(data, "num_graphs", None))
        prec = _PRECISIONS[self.precision]
        x = _LinearFn.apply(x, self.fc1.weight.t(), self.fc1.bias, prec)
        if len(self.cfg["head"]) == 1:
            return torch.tanh(x) if self.cfg["final"] == "tanh" else x
        return _LinearFn.apply(torch.relu(x), self.fc2.weight.t(), self.fc2.bias, prec)
