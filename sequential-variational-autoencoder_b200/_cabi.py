"""ctypes binding of libsvae.so (include/svae.h).  Thin by design: structs, prototypes, error mapping.

The library is built in-tree by ``__graft_entry__.build()`` (``make -C csrc``).  If it is missing, every entry point
raises - there is no Python / CPU fallback for the hot path."""
import ctypes as C
import os

MAX_LEVELS, MAX_STEPS, NAME_LEN = 8, 64, 128
OPERAND_FP32, OPERAND_BF16 = 0, 1
PF_THETA, PF_INERT, PF_DEAD, PF_XAVIER = 1, 2, 4, 8

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvae.so")


class SvaeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsvae error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32), ("levels", C.c_int32),
        ("latent_dims", C.c_int32 * MAX_LEVELS), ("filter_sizes", C.c_int32 * (MAX_LEVELS + 2)),
        ("mc_steps", C.c_int32), ("intermediate_reconstruction", C.c_int32), ("regularized_mask", C.c_uint64),
        ("first_step_loss_coeff", C.c_float), ("latent_mean_clip", C.c_float), ("prior_stddev", C.c_float),
        ("min_highway", C.c_float), ("max_highway", C.c_float), ("range_lo", C.c_float), ("range_hi", C.c_float),
        ("clip_value", C.c_float), ("adam_beta1", C.c_float), ("adam_beta2", C.c_float), ("adam_eps", C.c_float),
        ("max_batch", C.c_int32), ("train_capacity", C.c_int32), ("operand_dtype", C.c_int32),
        ("share_theta_weights", C.c_int32), ("share_phi_weights", C.c_int32), ("add_noise_to_chain", C.c_int32),
        ("reserved", C.c_int32 * 5), ("noise_stddevs", C.c_float * MAX_STEPS),
    ]


class Losses(C.Structure):
    _fields_ = [("total", C.c_float), ("final_recon", C.c_float), ("recon", C.c_float * MAX_STEPS),
                ("kl", C.c_float * MAX_STEPS)]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("ndim", C.c_int32), ("shape", C.c_int32 * 4), ("numel", C.c_int64),
                ("offset", C.c_int64), ("step", C.c_int32), ("flags", C.c_int32)]


class KernelStats(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("total_ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


_P = C.c_void_p
_F = C.POINTER(C.c_float)

# name -> (restype, argtypes): every symbol include/svae.h declares
PROTOTYPES = {
    "svae_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_P)]),
    "svae_destroy": (C.c_int, [_P]),
    "svae_last_error": (C.c_char_p, [_P]),
    "svae_version": (C.c_char_p, []),
    "svae_set_stream": (C.c_int, [_P, _P]),
    "svae_sync": (C.c_int, [_P]),
    "svae_param_count": (C.c_int, [_P]),
    "svae_param_table": (C.c_int, [C.POINTER(Config), C.POINTER(ParamInfo), C.c_int]),
    "svae_param_info_get": (C.c_int, [_P, C.c_int, C.POINTER(ParamInfo)]),
    "svae_param_slices": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "svae_param_set": (C.c_int, [_P, C.c_int, _P]),
    "svae_param_get": (C.c_int, [_P, C.c_int, _P]),
    "svae_grad_get": (C.c_int, [_P, C.c_int, _P]),
    "svae_adam_get": (C.c_int, [_P, C.c_int, _P, _P]),
    "svae_adam_set": (C.c_int, [_P, C.c_int, _P, _P]),
    "svae_adam_step_count": (C.c_int64, [_P]),
    "svae_adam_set_step_count": (C.c_int, [_P, C.c_int64]),
    "svae_param_arena": (_P, [_P]),
    "svae_grad_arena": (_P, [_P]),
    "svae_arena_numel": (C.c_int64, [_P]),
    "svae_arena_read": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P]),
    "svae_forward": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_uint64, C.c_float, _P, _P, _P]),
    "svae_backward": (C.c_int, [_P]),
    "svae_adam_step": (C.c_int, [_P, C.c_float]),
    "svae_train_step": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_uint64, C.c_float, C.c_float]),
    "svae_train_step_host": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_uint64, C.c_float, C.c_float, C.POINTER(Losses)]),
    "svae_apply_noise": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_uint64, _P]),
    "svae_apply_noise_host": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_uint64, _P]),
    "svae_train_step_host_denoise": (C.c_int, [_P, _P, C.c_int, _P, C.c_uint64, C.c_float, C.c_float, C.c_float,
                                               C.c_float, C.c_float, C.c_uint64, _P, C.POINTER(Losses)]),
    "svae_forward_host": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_uint64, C.c_float, _P, _P, _P, _P, C.POINTER(Losses)]),
    "svae_read_losses": (C.c_int, [_P, C.POINTER(Losses)]),
    "svae_generate": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P]),
    "svae_generate_host": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P]),
    "svae_set_chain_noise_host": (C.c_int, [_P, _P, C.c_int]),
    "svae_read_chain_samples_host": (C.c_int, [_P, _P, C.c_int]),
    "svae_nccl_unique_id": (C.c_int, [C.c_char_p, C.c_char_p]),
    "svae_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "svae_comm_destroy": (C.c_int, [_P]),
    "svae_launch_count": (C.c_int64, [_P]),
    "svae_activation_bytes": (C.c_int64, [_P]),
    "svae_tc_layers": (C.c_int, [_P]),
    "svae_profile_enable": (C.c_int, [_P, C.c_int]),
    "svae_profile_read": (C.c_int, [_P, C.POINTER(KernelStats), C.c_int]),
    "svae_op_conv2d": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int] * 7),
    "svae_op_conv2d_transpose": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int] * 7),
    "svae_op_conv2d_backward": (C.c_int, [_P, _P, _P, _P, _P, _P] + [C.c_int] * 7),
    "svae_op_conv2d_transpose_backward": (C.c_int, [_P, _P, _P, _P, _P, _P] + [C.c_int] * 7),
    "svae_op_fc": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 4),
    "svae_op_fc_backward": (C.c_int, [_P, _P, _P, _P, _P, _P] + [C.c_int] * 4),
    "svae_op_tc_supported": (C.c_int, [C.c_int] * 7),
    "svae_op_tc2_supported": (C.c_int, [C.c_int] * 8),
    "svae_debug_block_tensor": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int64, C.POINTER(C.c_int32)]),
    "svae_debug_set_buffer": (C.c_int, [_P]),
    "svae_op_bn_act": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int, C.c_int]),
    "svae_op_bn_act_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int64, C.c_int, C.c_int]),
    "svae_op_adam": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_int64, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_float]),
}

_lib = None


def lib():
    """Load libsvae.so (once).  Raises if it has not been built: the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvaeError(-2, "libsvae.so is not built (%s); run __graft_entry__.build() or `make -C %s`"
                            % (LIB_PATH, os.path.join(_HERE, "csrc")))
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(handle, rc):
    if rc != 0:
        msg = lib().svae_last_error(handle)
        raise SvaeError(rc, msg.decode() if msg else "unknown")
    return rc


def nccl_library_path():
    """Path of the NCCL PyTorch bundles (same build torch.distributed uses), or None to let dlopen search."""
    try:
        import nvidia.nccl  # type: ignore

        d = os.path.join(os.path.dirname(nvidia.nccl.__path__[0] if hasattr(nvidia.nccl, "__path__") else nvidia.nccl.__file__), "nccl", "lib")
        for cand in (os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2"), os.path.join(d, "libnccl.so.2")):
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None
