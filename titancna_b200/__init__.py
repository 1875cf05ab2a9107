"""titancna_b200 -- B200-native (sm_100a) implementation of TitanCNA's E-step hot path.

Host-side mirror of the reference's R API for the path (same function names):
    loadDefaultParameters, setupClonalParameters, runEMclonalCN, viterbiClonalCN
over the C ABI in include/titancna_b200.h (titancna_b200/csrc/libtitancna_b200.so).
"""
from .params import loadDefaultParameters, setupClonalParameters, estimateDirichletParamsMap, dataMoments  # noqa: F401
from .hmm import (Solver, runEMclonalCN, viterbiClonalCN, fwd_back_clonalCN, viterbi_clonalCN,  # noqa: F401
                  chromosome_offsets)
from .sweep import runEMclonalCN_sharded, run_sweep, lpt_assign, shard_chromosomes, solution_grid, make_comm  # noqa: F401
from .results import outputTitanResults, outputTitanSegments, removeEmptyClusters, correctIntegerCN, decoupleMegaVar, decodeLOH, getMajorMinorCN  # noqa: F401
from .selection import computeSDbwIndex, outputModelParameters, selectSolution  # noqa: F401
from .staging import loadAlleleCounts, filterData, removeCentromere, getPositionOverlap, setGenomeStyle  # noqa: F401
from ._lib import TitanError, device_count, lib, LIB_PATH, Comm, measure_fp64_peak  # noqa: F401

__version__ = "0.1.0"
