"""recbole-fairrec_b200: B200-native (sm_100a) hot path of RecBole-FairRec behind the reference's plugin API.

Only what the path needs lives here:
  csrc/            hand-written CUDA kernels + the C ABI (include/fairrec_b200.h) -> libfairrec_b200.so
  _lib.py          ctypes binding (fails loudly when the library is missing; no CPU fallback)
  kernels.py       tensor-level wrappers of the C ABI
  focf.py          FOCF model (calculate_loss / predict / full_sort_predict + fused train_step)
  ops.py/layers.py autograd Functions over the generic layer kernels; MLPLayers mirror
  pfcn.py          PFCN_MLP / PFCN_PMF / PFCN_BiasedMF / PFCN_DMF models + alternating PFCNTrainer
  fairgo.py        FairGo_PMF / FairGo_GCN models (GCN pretrain, filtered fine-tune) + FairGoTrainer (SpMM + full-table filters)
  nfcf.py          NFCF model (NCF tower + BCE + differential-fairness regulariser) + NFCFTrainer (both stages)
  dataloader.py    device-side FOCF batch builder (FOCFDataLoader)
  evaluator.py     fused full-sort fair evaluation (EvalData, FullSortEvaluator)
  sampled_eval.py  sampled-negative (uni100) ranking evaluation (SampledEvalData, SampledEvaluator)
  sharded.py       row-sharded FOCF training over NVLink peer memory (ShardedFOCF, ShardedGroupEmu)
  trainer.py       FOCFTrainer (fit / evaluate)
  atomic.py        atomic-file datasets (.inter/.user/.item) -> ids, splits, history/positive lists (reference-identical)
  quick_start.py   run_recbole(model, dataset, config_file_list, config_dict)
  utils.py         get_model / get_trainer (plugin discovery by name, utils/utils.py:51-94)
  synth.py         synthetic data of the benchmark shapes
The directory name carries a hyphen; import it as `recbole_fairrec_b200` (shim at the repo root).
"""
from . import _lib, kernels  # noqa: F401
from .atomic import AtomicDataset  # noqa: F401
from .config import Config  # noqa: F401
from .dataloader import FOCFDataLoader, TrainData  # noqa: F401
from .evaluator import EvalData, FullSortEvaluator  # noqa: F401
from .fairgo import FairGo_GCN, FairGo_GCNTrainer, FairGo_PMF, FairGo_PMFTrainer, FairGoTrainer  # noqa: F401
from .focf import FOCF, pack_host_batch  # noqa: F401
from .interaction import Interaction  # noqa: F401
from .nfcf import NFCF, NFCFTrainer  # noqa: F401
from .pfcn import (PFCN_MLP, PFCN_PMF, PFCN_BiasedMF, PFCN_DMF, PFCNTrainer, PFCN_MLPTrainer, PFCN_PMFTrainer,  # noqa: F401
                   PFCN_BiasedMFTrainer, PFCN_DMFTrainer)
from .sharded import ShardedFOCF, ShardedFOCFLoader, ShardedGroupEmu, ShardedTrainData  # noqa: F401
from .sampled_eval import SampledEvalData, SampledEvaluator, sample_negatives  # noqa: F401
from .trainer import FOCFTrainer  # noqa: F401
from .utils import get_model, get_trainer  # noqa: F401

__version__ = "0.1.0"
