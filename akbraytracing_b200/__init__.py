"""akbraytracing_b200 -- B200-native (sm_100a) drop-in for the two data-parallel hot paths of
Kakekakechan/AKBRaytracing: the Huygens-Fresnel pair sum (wavecalc) and the ray / quadric-mirror
chain (raytrace).  Hand-written CUDA behind a C-ABI (include/akb_b200.h); no CPU fallback."""
from . import _lib
from .wavecalc import (PHASE_EXACT, PHASE_FAITHFUL, PHASE_REFERENCED, WaveField3D, compute_u, compute_u_parallel,
                       forward_propagation_cupy_batch,
                       forward_propagation_cupy_batch_multi_gpu, forward_propagation_numpy_batch,
                       fresnel_sum, fresnel_sum_sharded)
from .raytrace import (PlanePoints, ell, intersect_reflect, mirr_ray_intersection, norm_vector,
                       normalize_vector, plane_ray_intersection, reflect_ray, rotation_matrices, trace_chain,
                       trace_chain_batched, wavefront_opl)
from .handoff import calc_dS, opl_to_field
from .psf import compute_psf_fft, field_to_pupil, psf_to_db
from .throughfocus import compute_psf_fft_batch, fresnel_sum_planes, psf_stack
from .stagechain import auto_phase_mode, load_handoff, parse_conditions, run_stage_chain, write_handoff

__version__ = "0.2.0"
