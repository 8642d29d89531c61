"""Install the scan2cap_b200 mirrors under the reference's own import names, so that the reference's callers
(scripts/train.py:25, benchmark/predict.py:28, lib/solver.py:21, models/*.py) run unchanged:

    import scan2cap_b200.dropin; scan2cap_b200.dropin.install()
    from models.capnet import CapNet                                   # -> scan2cap_b200.models.capnet.CapNet
    from lib.pointnet2.pointnet2_modules import PointnetSAModuleVotes  # -> the sm_100a-backed module
    import pointnet2._ext as _ext                                      # -> libs2c.so through its C ABI

Only the modules of the hot path are replaced; everything else of the reference (datasets, solver, evaluation)
keeps importing from the reference tree.
"""
import importlib
import sys
import types

_MAP = {
    "pointnet2._ext": "scan2cap_b200.lib.pointnet2._ext",
    "lib.pointnet2.pointnet2_utils": "scan2cap_b200.lib.pointnet2.pointnet2_utils",
    "lib.pointnet2.pointnet2_modules": "scan2cap_b200.lib.pointnet2.pointnet2_modules",
    "lib.pointnet2.pytorch_utils": "scan2cap_b200.lib.pointnet2.pytorch_utils",
    "pointnet2_utils": "scan2cap_b200.lib.pointnet2.pointnet2_utils",   # the reference also imports these flat,
    "pytorch_utils": "scan2cap_b200.lib.pointnet2.pytorch_utils",       # via its sys.path hack (pointnet2_modules.py:14-17)
    "models.backbone_module": "scan2cap_b200.models.backbone_module",
    "models.voting_module": "scan2cap_b200.models.voting_module",
    "models.proposal_module": "scan2cap_b200.models.proposal_module",
    "models.graph_module": "scan2cap_b200.models.graph_module",
    "models.caption_module": "scan2cap_b200.models.caption_module",
    "models.capnet": "scan2cap_b200.models.capnet",
    "models.capnet_pretrained": "scan2cap_b200.models.capnet_pretrained",
    "models.mask_votenet": "scan2cap_b200.models.mask_votenet",
    "models.encoder_module": "scan2cap_b200.models.encoder_module",
    "lib.loss_helper_pretrained": "scan2cap_b200.lib.loss_helper_pretrained",
    "lib.loss_helper": "scan2cap_b200.lib.loss_helper",
    "utils.nn_distance": "scan2cap_b200.utils.nn_distance",
}


def install(replace_loss=True):
    """Alias the mirrors into sys.modules.  Parent packages that are not importable yet (e.g. `pointnet2`,
    which the reference installs with lib/pointnet2/setup.py) are created as empty namespace modules."""
    for alias, target in _MAP.items():
        if not replace_loss and alias in ("lib.loss_helper", "lib.loss_helper_pretrained", "utils.nn_distance"):
            continue
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, child = alias.rpartition(".")
        if parent:
            if parent not in sys.modules:
                try:
                    importlib.import_module(parent)
                except Exception:
                    sys.modules[parent] = types.ModuleType(parent)
            setattr(sys.modules[parent], child, mod)
    return sorted(_MAP)
