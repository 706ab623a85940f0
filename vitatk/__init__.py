"""vitatk — B200-native white-box attack engine (import shim).

The sources live in ``adapting-pretrained-vision-transformers-with-lora-against-attack-vectors_b200/``
(a directory name Python cannot import directly); this package extends its ``__path__`` there so that
``vitatk.attacks``, ``vitatk.engine`` ... resolve to those files.
"""
import os as _os

PACKAGE_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "adapting-pretrained-vision-transformers-with-lora-against-attack-vectors_b200",
)
__path__.append(PACKAGE_DIR)

from .attacks import (  # noqa: E402,F401
    FGSM,
    EngineModule,
    PGD,
    LogitsModel,
    NormalizedModel,
    attack,
    batched_fgsm_attack,
    compile_model,
    get_model_output,
    invalidate,
)
from .engine import Engine  # noqa: E402,F401
from .patch import AdversarialPatch, sample_transforms  # noqa: E402,F401
from .swin import SwinEngine  # noqa: E402,F401
from .training import LoraTrainer, average_gradients, init_adapters  # noqa: E402,F401
from .adapters import (  # noqa: E402,F401
    PeftAdapter,
    compose,
    find_lora_adapters,
    load_engine,
    read_adapter,
    write_adapter,
)
