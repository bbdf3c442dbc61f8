"""TEST INFRASTRUCTURE: binds ``ReconstructInducer.func`` to the float64 oracle instead of the CUDA engine, so the host
logic of the driver (epoch loop, sampler call order, clustering, B-cubed, checkpoint) runs on the CPU test tier and so
that golden end-to-end traces can be generated.  Never imported by the product package."""
from __future__ import annotations

import numpy as np

from oracle import rae_oracle as O


def oracle_backend(inducer):
    data = inducer.data
    p = {k: np.asarray(v, dtype=np.float64) for k, v in inducer.initial_params.items()}
    acc0 = getattr(inducer, 'initial_acc', None)
    om = O.OracleModel(inducer.decoder_type, p, K=inducer.relationNum, d=inducer.embedSize, S=inducer.neg_sample_num,
                       B=inducer.batch_size, lr=inducer.learningRate, l1=inducer.lambdaL1, l2=inducer.lambdaL2,
                       alpha=inducer.alpha, optimizer=inducer.optimization, ext_reg=inducer.extendedReg)
    if acc0 is not None:                       # a resumed run (ReconstructInducer.load before compile_function)
        om.acc = {k: np.asarray(v, dtype=np.float64) for k, v in acc0.items()}
    func = {}
    for split in data.generate_split_keys():
        sp = data.split[split]
        om.bind_split(split, sp.indptr, sp.indices, sp.args1, sp.args2)
        func['label_' + split] = (lambda b, _s=split: om.label(_s, b))
    func['train'] = om.train
    inducer.oracle_model = om
    inducer.get_parameters = lambda: {k: v.copy() for k, v in om.params.items()}
    inducer.get_accumulators = lambda: {k: v.copy() for k, v in om.acc.items()}
    return func
