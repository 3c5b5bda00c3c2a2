"""Name -> class registry, the plug-in point every reference entry script uses
(``immunostruct/models/mapping.py:6-21``; consumed at train_IEDB_wFT.py:60-62,
train_Cancer_wFT.py:72-74, infer_IEDB_or_Cancer.py:59-62)."""
from .ablation_models import (DualModel, SequenceFpModel, SequenceModel, StructureModel, StructureModel_SSL,
                              StructureModelv2)
from .comparative_models import (HybridModel_Comparative, HybridModel_Comparative_SSL, HybridModelv2_Comparative,
                                 HybridModelv2_Comparative_SSL)
from .hybrid_models import HybridModel, HybridModel_SSL, HybridModelv2, HybridModelv2_SSL

model_map = {cls.__name__: cls for cls in (
    SequenceModel, SequenceFpModel, StructureModel, StructureModel_SSL, StructureModelv2,
    HybridModel, HybridModel_SSL, HybridModelv2, HybridModelv2_SSL,
    HybridModel_Comparative, HybridModel_Comparative_SSL, HybridModelv2_Comparative,
    HybridModelv2_Comparative_SSL, DualModel)}
