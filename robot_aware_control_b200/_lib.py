"""ctypes binding of libracb200.so (include/racb200.h). The library is built in-tree by
`python -m robot_aware_control_b200.build` (or __graft_entry__.build()). There is no fallback: a missing library is an
ImportError-like RuntimeError at first use, a failing call raises with rac_last_error()."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libracb200.so")

RAC_OK, RAC_ERR_INVALID, RAC_ERR_CUDA, RAC_ERR_STATE, RAC_ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class RacConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "image_height", "image_width", "g_dim", "z_dim", "action_dim", "robot_dim", "use_mask", "use_future_mask",
        "use_robot_state", "use_future_robot_state", "conv_impl", "lstm_group_norm")]


class RacStep(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("image", C.c_void_p), ("mask", C.c_void_p), ("robot", C.c_void_p),
        ("robot_next", C.c_void_p), ("action", C.c_void_p), ("eps", C.c_void_p), ("seed", C.c_ulonglong),
        ("noise_ctr", C.c_uint), ("sample_mean", C.c_int), ("use_posterior", C.c_int), ("next_robot", C.c_void_p),
        ("eps_post", C.c_void_p), ("force_use_prior", C.c_int), ("keep_skip", C.c_int), ("x_pred", C.c_void_p),
        ("mu_p", C.c_void_p), ("logvar_p", C.c_void_p), ("mu", C.c_void_p), ("logvar", C.c_void_p),
    ]


class RacTrainStep(C.Structure):
    _fields_ = [
        ("image", C.c_void_p), ("mask", C.c_void_p), ("robot", C.c_void_p), ("next_robot", C.c_void_p),
        ("action", C.c_void_p), ("eps_prior", C.c_void_p), ("eps_post", C.c_void_p), ("seed", C.c_ulonglong),
        ("noise_step", C.c_ulonglong), ("keep_skip", C.c_int), ("x_pred", C.c_void_p), ("mu", C.c_void_p),
        ("logvar", C.c_void_p), ("mu_p", C.c_void_p), ("logvar_p", C.c_void_p),
    ]


class RacRollout(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("steps", C.c_int), ("cand_offset", C.c_int), ("actions", C.c_void_p),
        ("start_img", C.c_void_p), ("goal_imgs", C.c_void_p), ("num_goals", C.c_int), ("goal_masks", C.c_void_p),
        ("states", C.c_void_p), ("state_t_stride", C.c_int64), ("masks", C.c_void_p), ("mask_t_stride", C.c_int64),
        ("eps", C.c_void_p), ("seed", C.c_ulonglong), ("noise_ctr_base", C.c_uint), ("sample_mean", C.c_int),
        ("zero_robot", C.c_int), ("dontcare_cost", C.c_int), ("sparse_cost", C.c_int),
        ("world_cost_weight", C.c_float), ("obs_out", C.c_void_p), ("step_cost_out", C.c_void_p),
        ("sum_cost", C.c_void_p), ("peer_cost_bufs", C.c_void_p), ("peer_world", C.c_int),
        ("peer_offset", C.c_int64),
    ]


class RacRobotModel(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("low", C.c_float * 5), ("high", C.c_float * 5), ("frame_diff", C.c_double * 2),
        ("push_height", C.c_double), ("cam_center", C.c_float * 3), ("cam_minv", C.c_float * 9),
        ("shoulder_z", C.c_float), ("l_upper", C.c_float), ("l_fore", C.c_float), ("l_wrist", C.c_float),
        ("pitch", C.c_float), ("radius", C.c_float * 4),
    ]


class RacCem(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("steps", C.c_int), ("iters", C.c_int), ("topk", C.c_int), ("init_std", C.c_float),
        ("clamp", C.c_float), ("std_floor", C.c_float), ("noise", C.c_void_p), ("rollout", RacRollout),
        ("robot", C.POINTER(RacRobotModel)), ("robot_start_state", C.c_void_p), ("robot_render_masks", C.c_int),
        ("robot_extra_radius", C.c_float),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "rac_abi_version": (C.c_int, []),
    "rac_create": (C.c_int, [C.POINTER(RacConfig), C.POINTER(C.c_void_p)]),
    "rac_destroy": (C.c_int, [C.c_void_p]),
    "rac_last_error": (C.c_char_p, [C.c_void_p]),
    "rac_layer_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rac_load_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "rac_load_lstm_norm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "rac_prepare": (C.c_int, [C.c_void_p, C.c_int]),
    "rac_init_hidden": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "rac_forward": (C.c_int, [C.c_void_p, C.POINTER(RacStep), C.c_void_p]),
    "rac_rollout_cost": (C.c_int, [C.c_void_p, C.POINTER(RacRollout), C.c_void_p]),
    "rac_peer_barrier": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p]),
    "rac_cem_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rac_topk": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rac_cem_refit": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "rac_cem_plan": (C.c_int, [C.c_void_p, C.POINTER(RacCem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "rac_masked_cost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.c_int, C.c_void_p]),
    "rac_l1_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "rac_dontcare_l1_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p]),
    "rac_recon_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rac_predict_states": (C.c_int, [C.POINTER(RacRobotModel), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "rac_render_masks": (C.c_int, [C.POINTER(RacRobotModel), C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "rac_robot_world_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rac_kl_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                              C.c_void_p]),
    "rac_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "rac_world_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rac_ssim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                           C.c_int, C.c_void_p]),
    "rac_composite": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rac_process_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "rac_preprocess_states": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "rac_debug_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int)]),
    "rac_launch_count": (C.c_int64, [C.c_void_p]),
    "rac_train_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "rac_train_destroy": (C.c_int, [C.c_void_p]),
    "rac_train_forward_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "rac_train_adam_step": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rac_train_set_adam_step": (C.c_int, [C.c_void_p, C.c_int]),
    "rac_train_set_grad_scale": (C.c_int, [C.c_void_p, C.c_float]),
    "rac_train_unpack_deferred": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rac_train_invalidate_packed": (C.c_int, [C.c_void_p]),
    "rac_train_debug_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "rac_train_step_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rac_train_step_forward": (C.c_int, [C.c_void_p, C.POINTER(RacTrainStep), C.c_void_p]),
    "rac_train_step_backward": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(RacTrainStep), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rac_profile_begin": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "rac_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Loads libracb200.so once. Raises if it has not been built -- there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension with `python -m robot_aware_control_b200.build` "
            "(robot_aware_control_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RacError(RuntimeError):
    pass


def check(code, handle=None, what=""):
    if code == RAC_OK:
        return
    msg = ""
    if handle:
        raw = load().rac_last_error(handle)
        msg = raw.decode("utf-8", "replace") if raw else ""
    text = f"{what} failed with rac_status {code}: {msg}"
    if code == RAC_ERR_INVALID:
        raise ValueError(text)
    if code == RAC_ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise RacError(text)


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor, or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
