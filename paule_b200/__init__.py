"""paule_b200 -- B200-native implementation of PAULE's gradient-planning hot path.

Public surface mirrors the reference package for that path:
  paule_b200.paule.Paule / PAULE, PlanningResults        (reference: paule/paule.py)
  paule_b200.models.ForwardModel, EmbeddingModel, InverseModelMelTimeSmoothResidual (+ aliases)
  paule_b200.planner.BatchPlanner                          (the batched, device-resident inner loop)
  paule_b200.distributed                                   (word sharding over GPUs + final gather)
All compute runs in libpaule_b200.so (hand-written sm_100a CUDA behind a C ABI, include/paule_b200.h).
"""
from . import _lib, ops, models, planner, branches, paule, distributed  # noqa: F401
from .paule import Paule, PAULE, PlanningResults  # noqa: F401
from .models import (ForwardModel, EmbeddingModel, InverseModelMelTimeSmoothResidual,  # noqa: F401
                     InverseModel, MelEmbeddingModel, MelEmbeddingModelMelSmoothResidualUpsampling,
                     LinearClassifier, Generator)
from .planner import BatchPlanner  # noqa: F401

__version__ = "0.1.0"
