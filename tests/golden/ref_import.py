"""Import the UNMODIFIED reference (XHwind/HM-ViT) from /root/reference on CPU.

Only usable in the build container (the GPU box has no /root/reference).  The
reference imports matplotlib / shapely / mmdet3d at module scope although the
fusion path never uses them; they are absent here, so empty stand-ins are put
in sys.modules before the import.  Nothing of the reference is
copied: it is imported where it lies.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _find_root() -> str:
    """the reference installed under baseline/_ref/ (pip --target, git-ignored, travels to the GPU
    box), then /root/reference (build container only); HMVIT_REFERENCE overrides."""
    cands = [os.environ.get("HMVIT_REFERENCE"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference"]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "opencood")):
            return c
    return cands[-1]


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "opencood"))


def load():
    """Returns a namespace with the reference classes/functions on the hot path."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "shapely", "shapely.geometry"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "shapely.geometry":
                    m.Polygon = object
                sys.modules[name] = m
    if "opencood.models.bevformer_wrapper" not in sys.modules:
        fake = types.ModuleType("opencood.models.bevformer_wrapper")
        fake.BEVFormerWrapper = object
        sys.modules["opencood.models.bevformer_wrapper"] = fake
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    from opencood.models.bevformer_point_pillar_hetero import HeteroFusion
    from opencood.models.sub_modules.hetero_fusion import HeteroFusionBlock, HeteroAttention
    from opencood.models.sub_modules.fuse_utils import regroup
    from opencood.models.sub_modules.spatial_transformation import SpatialTransformation
    from opencood.models.sub_modules.torch_transformation_utils import get_roi_and_cav_mask
    from opencood.models.sub_modules.hetero_decoder import HeteroDecoder
    ns = types.SimpleNamespace(HeteroFusion=HeteroFusion, HeteroFusionBlock=HeteroFusionBlock,
                               HeteroAttention=HeteroAttention, regroup=regroup,
                               SpatialTransformation=SpatialTransformation,
                               get_roi_and_cav_mask=get_roi_and_cav_mask, HeteroDecoder=HeteroDecoder)
    return ns
