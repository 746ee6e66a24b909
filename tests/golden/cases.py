"""Seeded chain builders shared by make_golden.py and the tests (so that the tests can rebuild the exact chains
whose packed weights are stored in the golden files)."""
import numpy as np

from oracle import dflow_oracle as O


def readme_n1_chain(x_fixture):
    """test/runtests.jl:104-109: masks [1,2,3],[3,4,5],[5,1,2], hidden 16, NormalizationLayer(x, -1, 1); n = 1."""
    return O.readme_chain(1, x_fixture, seed=42)


def ref_chain_d7(xs):
    """test/runtests.jl:66-95: layer [1,3,5,7], layer [4,2,5,1,6] (unsorted), CouplingBlock [4,2,5,1], Normalization."""
    r = np.random.default_rng(3)
    l1 = O.coupling_layer(O.coupling_axes(7, [1, 3, 5, 7], n=2), rng=r, bias_scale=0.1)
    l2 = O.coupling_layer(O.coupling_axes(7, [4, 2, 5, 1, 6], n=2), rng=r, bias_scale=0.1)
    blk = O.coupling_block(O.coupling_axes(7, [4, 2, 5, 1], n=2), rng=r, bias_scale=0.1)
    return O.concatenate((O.Chain([l1, l2]), O.Chain([blk, O.norm_layer_from_data(xs)])))
