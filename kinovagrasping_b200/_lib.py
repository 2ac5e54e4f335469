"""ctypes binding of the C ABI in include/kg.h.

The package loads exactly one library: kinovagrasping_b200/libkg_b200.so, the nvcc-built sm_100a
extension.  There is no CPU fallback: if the library is missing or no CUDA device answers, loading
raises.  (tests/ reuse `bind` to put the same signatures on their host-emulation build of the same
source; that library is a debugging aid and is never loaded from here.)"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkg_b200.so")
DEBUG_FLOATS = 2020

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

SIGNATURES = {
    "kg_create": (C.c_int, [C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(_vp)]),
    "kg_destroy": (C.c_int, [_vp]),
    "kg_last_error": (C.c_char_p, [_vp]),
    "kg_num_envs": (C.c_int, [_vp]),
    "kg_set_stream": (C.c_int, [_vp, _vp]),
    "kg_sync": (C.c_int, [_vp]),
    "kg_reset": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "kg_reset_sampled": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _vp]),
    "kg_step": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "kg_step_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "kg_expert_action": (C.c_int, [_vp, C.c_int, _vp]),
    "kg_expert_action_host": (C.c_int, [_vp, C.c_int, _vp]),
    "kg_reset_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "kg_reset_pose": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kg_reset_pose_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "kg_set_hand_location": (C.c_int, [_vp, C.c_int]),
    "kg_set_shape_cycle": (C.c_int, [_vp, _vp, C.c_int]),
    "kg_get_orientation_host": (C.c_int, [_vp, _vp]),
    "kg_set_action_dim": (C.c_int, [_vp, C.c_int]),
    "kg_get_state": (C.c_int, [_vp, _vp, _vp]),
    "kg_set_state": (C.c_int, [_vp, _vp, _vp]),
    "kg_get_state_host": (C.c_int, [_vp, _vp, _vp]),
    "kg_set_state_host": (C.c_int, [_vp, _vp, _vp]),
    "kg_set_max_episode_steps": (C.c_int, [_vp, C.c_int]),
    "kg_set_frame_skip": (C.c_int, [_vp, C.c_int]),
    "kg_set_env_offset": (C.c_int, [_vp, C.c_int]),
    "kg_set_schedule": (C.c_int, [_vp, C.c_int]),
    "kg_set_mode": (C.c_int, [_vp, C.c_int]),
    "kg_set_obs_global": (C.c_int, [_vp, _vp]),
    "kg_obs_global_buffer": (C.c_int, [_vp, C.POINTER(_vp)]),
    "kg_get_obs_global_host": (C.c_int, [_vp, _vp]),
    "kg_set_pipeline": (C.c_int, [_vp, C.c_int]),
    "kg_measure_fp32_peak": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "kg_set_timing": (C.c_int, [_vp, C.c_int]),
    "kg_set_graph": (C.c_int, [_vp, C.c_int]),
    "kg_set_dependent_launch": (C.c_int, [_vp, C.c_int]),
    "kg_timing_breakdown": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "kg_get_last_action_host": (C.c_int, [_vp, _vp]),
    "kg_set_grasp_forest": (C.c_int, [_vp, _vp, C.c_size_t]),
    "kg_set_grasp_reward": (C.c_int, [_vp, C.c_int]),
    "kg_get_grasp_reward_host": (C.c_int, [_vp, _vp, _vp]),
    "kg_grasp_predict_host": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "kg_get_stats": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "kg_get_prof": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "kg_forward_debug": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "kg_last_step_ms": (C.c_int, [_vp, _fp]),
}


def bind(cdll):
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)      # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def load():
    """Load the CUDA extension; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). kinovagrasping_b200 has no CPU fallback.")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib
