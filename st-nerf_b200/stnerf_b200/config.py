"""cfg stub for hosts without yacs: the fields `LayeredRFRender.__init__` reads (modeling/layered_rfrender.py:23-37),
with the shipped configs' values (configs/config_taekwondo.yml:51-66), plus the two B200 knobs."""
import types


def make_cfg(layer_num, n1, n2, use_space_time, precision="exact", chunk_rays=0):
    M = types.SimpleNamespace(
        BOARDER_WEIGHT=1e10, SAMPLE_METHOD="BBOX", SAME_SPACENET=False, TKERNEL_INC_RAW=True,
        POSE_REFINEMENT=False, USE_DIR=True, USE_DEFORM_VIEW=False, USE_DEFORM_TIME=True,
        USE_SPACE_TIME=use_space_time, BKGD_USE_DEFORM_TIME=False, BKGD_USE_SPACE_TIME=False,
        DEEP_RGB=False, COARSE_RAY_SAMPLING=n1, FINE_RAY_SAMPLING=n2, B200_PRECISION=precision,
        B200_CHUNK_RAYS=chunk_rays)
    return types.SimpleNamespace(MODEL=M, DATASETS=types.SimpleNamespace(LAYER_NUM=layer_num))
