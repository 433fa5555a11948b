"""ctypes mirror of include/g4hepem_b200.h (struct layouts + function prototypes).

The library is loaded lazily; a missing extension is a hard error (there is no CPU fallback).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# G4HB200_LIB: an alternative build of the same library (kernel tuning A/B runs)
LIB_PATH = os.environ.get("G4HB200_LIB") or os.path.join(_HERE, "csrc", "libg4hepem_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class ElectronTables(C.Structure):
    _fields_ = [
        ("num_loss", C.c_int32),
        ("loss_log_min_ekin", C.c_double),
        ("loss_eil_delta", C.c_double),
        ("loss_egrid", c_dp),
        ("loss_data", c_dp),
        ("resmx_start", c_ip),
        ("resmx_data", c_dp),
        ("num_resmx", C.c_int32),
        ("enuc_log_min_ekin", C.c_double),
        ("enuc_eil_delta", C.c_double),
        ("enuc_egrid", c_dp),
        ("enuc_data", c_dp),
        ("tr1_data", c_dp),
        ("sel_ioni_start", c_ip),
        ("sel_ioni_data", c_dp),
        ("num_sel_ioni", C.c_int32),
        ("sel_sb_start", c_ip),
        ("sel_sb_data", c_dp),
        ("num_sel_sb", C.c_int32),
        ("sel_rb_start", c_ip),
        ("sel_rb_data", c_dp),
        ("num_sel_rb", C.c_int32),
    ]


class Tables(C.Structure):
    _fields_ = [
        ("electron_tracking_cut", C.c_double),
        ("gamma_tracking_cut", C.c_double),
        ("min_loss_table_energy", C.c_double),
        ("electron_brem_model_lim", C.c_double),
        ("is_msc_positron_cor", C.c_int32),
        ("is_msc_displacement", C.c_int32),
        ("num_regions", C.c_int32),
        ("region_pars", c_dp),
        ("num_matcut", C.c_int32),
        ("mc_cuts", c_dp),
        ("mc_imat", c_ip),
        ("mc_ireg", c_ip),
        ("num_mat", C.c_int32),
        ("mat_num_elem", c_ip),
        ("mat_elem_start", c_ip),
        ("mat_elem_z", c_ip),
        ("mat_elem_natoms", c_dp),
        ("mat_pars", c_dp),
        ("mat_sandia_num", c_ip),
        ("mat_sandia_start", c_ip),
        ("elem_pars", c_dp),
        ("elem_sandia_num", c_ip),
        ("elem_sandia_start", c_ip),
        ("num_sandia", C.c_int32),
        ("sandia_energies", c_dp),
        ("sandia_cof", c_dp),
        ("electron", ElectronTables),
        ("positron", ElectronTables),
        ("sb_log_min_el_energy", C.c_double),
        ("sb_il_delta_el_energy", C.c_double),
        ("sb_el_energy", c_dp),
        ("sb_lel_energy", c_dp),
        ("sb_lkappa", c_dp),
        ("sb_gcut_start", c_ip),
        ("sb_gcut_indices", c_ip),
        ("num_sb_gcut", C.c_int32),
        ("sb_start_per_z", c_ip),
        ("sb_data", c_dp),
        ("num_sb_data", C.c_int32),
        ("gm_data_per_mat", C.c_int32),
        ("gm_num_data0", C.c_int32),
        ("gm_num_data1", C.c_int32),
        ("gm_emax0", C.c_double),
        ("gm_log_emin0", C.c_double),
        ("gm_eil_delta0", C.c_double),
        ("gm_emax1", C.c_double),
        ("gm_log_emin1", C.c_double),
        ("gm_eil_delta1", C.c_double),
        ("gm_log_emin2", C.c_double),
        ("gm_eil_delta2", C.c_double),
        ("gm_mxsec", c_dp),
        ("gm_conv_egrid_size", C.c_int32),
        ("gm_conv_log_min_ekin", C.c_double),
        ("gm_conv_eil_delta", C.c_double),
        ("gm_conv_start", c_ip),
        ("gm_conv_egrid", c_dp),
        ("gm_conv_data", c_dp),
        ("num_gm_conv", C.c_int32),
    ]


ELECTRON_PAIR_GROUPS = (
    "ekin_logekin", "dirx_diry", "dirz_safety", "nia01", "nia23", "msc_irange_dynrf", "msc_tlimmin_gauss",
)
ELECTRON_RESULT_GROUPS = ("gstep_pstep", "edep_dispx", "dispy_dispz")
ELECTRON_HANDOVER_GROUPS = ("mfp01", "mfp23", "range_lambtr1", "tstep_zpath", "par12", "par3_pad", "prestep")


class ElectronBatch(C.Structure):
    _fields_ = (
        [("n", C.c_int64)]
        + [(g, c_dp) for g in ELECTRON_PAIR_GROUPS]
        + [("meta", c_ip)]
        + [(g, c_dp) for g in ELECTRON_RESULT_GROUPS]
        + [("winner", c_ip)]
        + [(g, c_dp) for g in ELECTRON_HANDOVER_GROUPS]
    )


GAMMA_PAIR_GROUPS = ("ekin_logekin", "dirx_diry", "dirz_nia0")
GAMMA_RESULT_GROUPS = ("gstep_mfp0", "edep_pemxsec")


class GammaBatch(C.Structure):
    _fields_ = (
        [("n", C.c_int64)]
        + [(g, c_dp) for g in GAMMA_PAIR_GROUPS]
        + [("meta", c_ip)]
        + [(g, c_dp) for g in GAMMA_RESULT_GROUPS]
        + [("winner", c_ip)]
    )


class SecondaryQueue(C.Structure):
    _fields_ = [
        ("capacity", C.c_int64),
        ("dirx_diry", c_dp),
        ("dirz_ekin", c_dp),
        ("parent_kind", c_ip),
        ("parent_slot", c_ip),
        ("count", c_ip),
        ("parent_base", C.c_int32),
    ]


class SlabGeometry(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32),
        ("num_absorbers", C.c_int32),
        ("absorber_thickness", C.c_double * 4),
        ("absorber_couple", C.c_int32 * 4),
        ("half_yz", C.c_double),
        ("woodcock_on", C.c_int32),
        ("woodcock_couple", C.c_int32),
        ("woodcock_ekin_min", C.c_double),
    ]


class ShowerStats(C.Structure):
    _fields_ = [
        ("num_steps", C.c_int64),
        ("electron_track_steps", C.c_int64),
        ("gamma_track_steps", C.c_int64),
        ("secondaries", C.c_int64),
        ("peak_electrons", C.c_int64),
        ("peak_gammas", C.c_int64),
        ("leak_electron", C.c_double),
        ("leak_gamma", C.c_double),
        ("device_ms", C.c_double),
        ("kernel_launches", C.c_int64),
        ("remaining_electrons", C.c_int64),
        ("remaining_gammas", C.c_int64),
        ("remaining_ekin", C.c_double),
    ]


F_POSITRON = 0x01
F_ON_BOUNDARY = 0x02
F_MSC_FIRST_STEP = 0x04
F_MSC_ACTIVE = 0x08
F_MSC_DISPLACE = 0x10
F_MSC_NO_SCATTER = 0x20
F_GAUSS_CACHED = 0x40
F_WDT_ON = 0x80

SEC_ELECTRON, SEC_POSITRON, SEC_GAMMA = 0, 1, 2
# op codes of g4hb200_electron_track_op / g4hb200_gamma_track_op (include/g4hepem_b200.h)
(OP_HOWFAR_DISCRETE, OP_HOWFAR_MSC, OP_UPDATE_PSTEP, OP_UPDATE_NIA, OP_MEAN_ELOSS, OP_SAMPLE_MSC, OP_LOSS_FLUCT, OP_DISCRETE,
 OP_ANNIHILATE_AT_REST, OP_PERFORM_CONTINUOUS, OP_RESAMPLE_NIA) = range(11)
GOP_HOWFAR_TRACK, GOP_UPDATE_NIA, GOP_SELECT_INTERACTION, GOP_PERFORM_SELECTED = range(4)
NUM_STAGES = 20

# every symbol include/g4hepem_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_H = C.c_void_p
PROTOTYPES = {
    "g4hb200_create": (C.c_int, [C.POINTER(Tables), C.c_int, C.POINTER(_H)]),
    "g4hb200_destroy": (C.c_int, [_H]),
    "g4hb200_last_error": (C.c_char_p, []),
    "g4hb200_device_count": (C.c_int, []),
    "g4hb200_electron_batch_alloc": (C.c_int, [_H, C.c_int64, C.POINTER(ElectronBatch)]),
    "g4hb200_electron_batch_free": (C.c_int, [_H, C.POINTER(ElectronBatch)]),
    "g4hb200_gamma_batch_alloc": (C.c_int, [_H, C.c_int64, C.POINTER(GammaBatch)]),
    "g4hb200_gamma_batch_free": (C.c_int, [_H, C.POINTER(GammaBatch)]),
    "g4hb200_secondary_queue_alloc": (C.c_int, [_H, C.c_int64, C.POINTER(SecondaryQueue)]),
    "g4hb200_secondary_queue_free": (C.c_int, [_H, C.POINTER(SecondaryQueue)]),
    "g4hb200_secondary_queue_reset": (C.c_int, [_H, C.POINTER(SecondaryQueue), _vp]),
    "g4hb200_electron_batch_upload": (C.c_int, [_H, C.POINTER(ElectronBatch), C.POINTER(ElectronBatch), _vp]),
    "g4hb200_electron_batch_download": (C.c_int, [_H, C.POINTER(ElectronBatch), C.POINTER(ElectronBatch), _vp]),
    "g4hb200_gamma_batch_upload": (C.c_int, [_H, C.POINTER(GammaBatch), C.POINTER(GammaBatch), _vp]),
    "g4hb200_gamma_batch_download": (C.c_int, [_H, C.POINTER(GammaBatch), C.POINTER(GammaBatch), _vp]),
    "g4hb200_secondary_queue_download": (C.c_int, [_H, C.POINTER(SecondaryQueue), C.POINTER(SecondaryQueue), _vp]),
    "g4hb200_sync": (C.c_int, [_H, _vp]),
    "g4hb200_device_alloc": (C.c_int, [_H, C.c_size_t, C.POINTER(_vp)]),
    "g4hb200_device_free": (C.c_int, [_H, _vp]),
    "g4hb200_memcpy": (C.c_int, [_H, _vp, _vp, C.c_size_t, C.c_int, _vp]),
    "g4hb200_electron_lookups": (C.c_int, [_H, C.c_int64, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "g4hb200_electron_lookups_f32": (C.c_int, [_H, C.c_int64, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "g4hb200_electron_stepping_xsecs": (C.c_int, [_H, C.c_int64, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "g4hb200_gamma_lookups": (C.c_int, [_H, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4hb200_select_target_element": (C.c_int, [_H, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4hb200_vdt_log_exp": (C.c_int, [_H, C.c_int64, _vp, _vp, _vp, _vp]),
    "g4hb200_rng_uniforms": (C.c_int, [_H, C.c_uint64, C.c_int64, _vp, C.c_int32, _vp, _vp]),
    "g4hb200_electron_howfar": (C.c_int, [_H, C.POINTER(ElectronBatch), C.c_uint64, _vp]),
    "g4hb200_electron_perform": (C.c_int, [_H, C.POINTER(ElectronBatch), C.POINTER(SecondaryQueue), C.c_uint64, _vp]),
    "g4hb200_electron_step": (C.c_int, [_H, C.POINTER(ElectronBatch), C.POINTER(SecondaryQueue), C.c_uint64, _vp]),
    "g4hb200_gamma_howfar": (C.c_int, [_H, C.POINTER(GammaBatch), C.c_uint64, _vp]),
    "g4hb200_gamma_perform": (C.c_int, [_H, C.POINTER(GammaBatch), C.POINTER(SecondaryQueue), C.c_uint64, _vp]),
    "g4hb200_gamma_step": (C.c_int, [_H, C.POINTER(GammaBatch), C.POINTER(SecondaryQueue), C.c_uint64, _vp]),
    "g4hb200_electron_track_op": (C.c_int, [_H, C.c_int, C.POINTER(ElectronBatch), C.POINTER(SecondaryQueue), C.c_uint64, _vp, _vp]),
    "g4hb200_electron_check_delta": (C.c_int, [_H, C.POINTER(ElectronBatch), _vp, _vp, _vp]),
    "g4hb200_gamma_track_op": (C.c_int, [_H, C.c_int, C.POINTER(GammaBatch), C.POINTER(SecondaryQueue), C.c_uint64, _vp]),
    "g4hb200_electron_step_host": (C.c_int, [_H, C.POINTER(ElectronBatch), C.POINTER(SecondaryQueue), C.c_uint64]),
    "g4hb200_gamma_step_host": (C.c_int, [_H, C.POINTER(GammaBatch), C.POINTER(SecondaryQueue), C.c_uint64]),
    "g4hb200_launch_count": (C.c_int64, [_H]),
    "g4hb200_set_kernel_timing": (C.c_int, [_H, C.c_int]),
    "g4hb200_set_msc_precision": (C.c_int, [_H, C.c_int]),
    "g4hb200_kernel_times": (C.c_int, [_H, _vp, _vp, _vp]),
    "g4hb200_stage_name": (C.c_char_p, [C.c_int]),
    "g4hb200_shower_run": (C.c_int, [_H, C.POINTER(SlabGeometry), C.c_int64, C.c_int32, C.c_double, C.c_uint64, C.c_int32,
                                     C.c_int64, C.c_int32, _vp, C.POINTER(ShowerStats)]),
    "g4hb200_mixed_run": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_uint64, C.c_int64, C.c_int32, _vp,
                                    C.POINTER(ShowerStats)]),
}

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(g4hepem_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class G4HB200Error(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load_library().g4hb200_last_error()
        raise G4HB200Error(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
