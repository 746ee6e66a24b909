"""densityflows.jl_b200 -- B200-native drop-in for the DensityFlows.jl coupling-chain hot path.

The directory name contains a dot, so import it through the `densityflows` shim at the repo root:

    import densityflows.jl_b200 as df

Public names mirror the reference's exports (src/DensityFlows.jl, src/Flows.jl:30-32, src/Chains.jl:27-28,
src/Data.jl:25-30); Julia's `f!` becomes `f_` (`train_`, `forward_`).
"""
from . import _lib
from ._lib import DflowError, DflowInvalidArg, DflowUnsupported
from .arrays import jl_empty, jl_full, jl_zeros, to_jl, to_numpy
from .data import (DataArrays, DataPartition, MetaData, device_permutation, dflt_theta, dflt_θ, maximum_θ, minimum_θ, normalize_input,
                   normalized_training_data, normalized_validation_data, number_conditions, number_dimensions,
                   resize_output, testing_data, training_data, validation_data)
from .flows import (Adam, Flow, LocalDataParallel, OptimiserState, PeerTrainStep, TrainStep, make_train_step, logpdf, pdf, predict, sample, sample_with_rejection, setup, train_, training_loss,
                    validation_loss)
from .model import (Chain, CouplingAxes, CouplingBlock, CouplingLayer, CouplingLayerBase, Dense, FlowChain,
                    FlowElement, NICECouplingLayer, NormalizationElement, NormalizationLayer, PackedChain,
                    RNVPCouplingLayer, backward, concatenate, forward, forward_, identity, is_reverse, minmax_rows,
                    relu, reverse, seed, sigmoid, tanh)

from .persist import load_flow, packed_parameters, save_flow

train_bang = train_
forward_bang = forward_
__version__ = "0.1.0"
