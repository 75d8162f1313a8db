"""ctypes binding of libdsr_b200.so (the C ABI declared in include/dsr_b200.h).

The library is the product: there is no CPU / eager fallback.  Importing this module when the
shared object has not been built raises ImportError with the build command; calling a compute
entry point without a CUDA device raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSR_B200_LIB: another build of the same library (diagnostics builds such as `make kstamp`)
LIB_PATH = os.environ.get('DSR_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'libdsr_b200.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing: build it with `make -C deep-super-resolution_b200/csrc` '
        '(or python -c "import __graft_entry__ as g; g.build()"); dsr_b200 has no CPU fallback.')

lib = C.CDLL(LIB_PATH)

vp, i32, i64, f32, u64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong, C.c_size_t


class StepBuffers(C.Structure):
    """dsr_step_buffers_t"""
    _fields_ = [(n, vp) for n in ('params', 'grads', 'adam_m', 'adam_v', 'bn_buffers', 'z_saved', 'z', 'lr_image',
                                  'out_hr', 'out_lr', 'g_out_lr', 'g_out_hr', 'loss_out')]


# name -> (restype, argtypes); every symbol of include/dsr_b200.h (tests/test_abi.py checks the two lists agree)
SIGNATURES = {
    'dsr_abi_version': (i32, []),
    'dsr_error_string': (C.c_char_p, [i32]),
    'dsr_lanczos_kernel': (i32, [i32, i32, C.POINTER(C.c_double), i32]),
    'dsr_downsampler_table_bytes': (sz, [i32, i32, i32, i32]),
    'dsr_downsampler_create': (i32, [C.POINTER(vp), i32, i32, i32, i32, vp, sz, vp]),
    'dsr_downsampler_destroy': (None, [vp]),
    'dsr_downsample_fwd': (i32, [vp, vp, vp, i32, vp]),
    'dsr_downsample_bwd': (i32, [vp, vp, vp, i32, vp]),
    'dsr_downsample_mse': (i32, [vp, vp, vp, vp, vp, vp, i32, vp]),
    'dsr_plan_create': (i32, [C.POINTER(vp), i32, i32, i32, i32, i32]),
    'dsr_plan_create_ex': (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, i32]),
    'dsr_plan_destroy': (None, [vp]),
    'dsr_plan_num_params': (i32, [vp]),
    'dsr_plan_param_numel': (i64, [vp]),
    'dsr_plan_param_info': (i32, [vp, i32, C.c_char_p, i32, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
    'dsr_plan_num_bn': (i32, [vp]),
    'dsr_plan_bn_numel': (i64, [vp]),
    'dsr_plan_bn_info': (i32, [vp, i32, C.c_char_p, i32, C.POINTER(i64), C.POINTER(i32)]),
    'dsr_plan_workspace_bytes': (sz, [vp]),
    'dsr_plan_bind': (i32, [vp, vp, sz, vp]),
    'dsr_net_forward': (i32, [vp, vp, vp, vp, vp, vp]),
    'dsr_net_backward': (i32, [vp, vp, vp, vp, vp, vp]),
    'dsr_adam_step': (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, vp]),
    'dsr_perturb': (i32, [vp, vp, i64, f32, u64, u64, vp]),
    'dsr_dip_step': (i32, [vp, vp, C.POINTER(StepBuffers), f32, f32, u64, i32, vp]),
    'dsr_dip_run': (i32, [vp, vp, C.POINTER(StepBuffers), f32, f32, u64, i32, i32, vp]),
    'dsr_plan_tensor': (i32, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                              C.POINTER(i32), C.POINTER(i32)]),
    'dsr_plan_last_launches': (i32, [vp]),
    'dsr_plan_set_debug_conv': (i32, [vp, i32]),
    'dsr_plan_debug_replay': (i32, [vp, C.c_char_p, i32, i32, vp]),
    'dsr_plan_set_profile': (i32, [vp, i32]),
    'dsr_plan_profile_read': (i32, [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i32)]),
    'dsr_plan_profile_top': (i32, [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    'dsr_plan_profile_dump': (i32, [vp, C.c_char_p, sz]),
    'dsr_plan_device_error': (i32, [vp, C.POINTER(i32)]),
    'dsr_debug_copy': (i32, [vp, vp, sz, vp]),
    'dsr_timeline_dump': (i32, [C.c_char_p, sz]),
    'dsr_gen_plan_create': (i32, [C.POINTER(vp), i32, i32, i32, i32, i32]),
    'dsr_gen_plan_destroy': (None, [vp]),
    'dsr_gen_state_numel': (i64, [vp]),
    'dsr_gen_num_tensors': (i32, [vp]),
    'dsr_gen_tensor_info': (i32, [vp, i32, C.c_char_p, i32, C.POINTER(i64), C.POINTER(i64)]),
    'dsr_gen_workspace_bytes': (sz, [vp]),
    'dsr_gen_bind': (i32, [vp, vp, sz, vp]),
    'dsr_gen_load_weights': (i32, [vp, vp, vp]),
    'dsr_gen_forward': (i32, [vp, vp, vp, vp]),
    'dsr_gen_last_launches': (i32, [vp]),
    'dsr_gen_debug_tensor': (i32, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                   C.POINTER(i32)]),
    'dsr_gen_device_error': (i32, [vp, C.POINTER(i32)]),
    'dsr_metric_workspace_bytes': (sz, []),
    'dsr_psnr': (i32, [vp, vp, i64, f32, vp, vp, vp]),
    'dsr_ssim': (i32, [vp, vp, i32, i32, i32, f32, vp, vp, vp]),
    'dsr_image_to_u8_hwc': (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    'dsr_plan_deterministic': (i32, [vp]),
    'dsr_gant_create': (i32, [C.POINTER(vp), i32, i32, i32, i32, i32, i32]),
    'dsr_gant_destroy': (None, [vp]),
    'dsr_gant_param_numel': (i64, [vp, i32]),
    'dsr_gant_buffer_numel': (i64, [vp, i32]),
    'dsr_gant_num_params': (i32, [vp, i32]),
    'dsr_gant_num_buffers': (i32, [vp, i32]),
    'dsr_gant_param_info': (i32, [vp, i32, i32, C.c_char_p, i32, C.POINTER(i64), C.POINTER(i64)]),
    'dsr_gant_buffer_info': (i32, [vp, i32, i32, C.c_char_p, i32, C.POINTER(i64), C.POINTER(i64)]),
    'dsr_gant_workspace_bytes': (sz, [vp]),
    'dsr_gant_bind': (i32, [vp, vp, sz, vp]),
    'dsr_gant_pack': (i32, [vp, i32, vp, vp]),
    'dsr_gant_g_forward': (i32, [vp, vp, vp, vp, vp, i32, vp]),
    'dsr_gant_g_backward': (i32, [vp, vp, vp, vp, vp]),
    'dsr_gant_d_forward': (i32, [vp, i32, vp, vp, vp, vp, vp]),
    'dsr_gant_d_backward': (i32, [vp, i32, vp, vp, f32, vp, vp]),
    'dsr_gant_d_backward_pair': (i32, [vp, vp, f32, f32, vp, vp]),
    'dsr_gant_bce': (i32, [vp, vp, f32, i32, vp, i32, vp]),
    'dsr_gant_dense_grad_overwrite': (i32, [vp, i32]),
    'dsr_gant_vgg_loss': (i32, [vp, vp, vp, vp, i32, vp, vp]),
    'dsr_gant_vgg_real': (i32, [vp, vp, vp]),
    'dsr_gant_device_error': (i32, [vp, C.POINTER(i32)]),
    'dsr_gant_last_launches': (i32, [vp]),
    'dsr_gant_tensor': (i32, [vp, C.c_char_p, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                             C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the .so is stale; rebuild
    _fn.restype = _res
    _fn.argtypes = _args


class DsrError(RuntimeError):
    pass


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = lib.dsr_error_string(rc).decode()
        raise DsrError(f'libdsr_b200: {what or "call"} failed with code {rc}: {msg}')


def require_cuda() -> None:
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('dsr_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
