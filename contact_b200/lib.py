"""ctypes loader of libcontact_addon_b200.so with the prototypes of include/contact_addon_b200.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libcontact_addon_b200.so")
_dll = None

ip, dp, cp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_char_p
I, D, L, V = C.c_int, C.c_double, C.c_long, C.c_void_p


class LibraryMissing(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/contact_addon_b200.h
PROTOTYPES = {
    "cntc_initializefirst": (None, [ip, ip, ip, cp, cp, cp, ip, ip, ip]),
    "cntc_initializefirst_new": (None, [ip, ip, ip, cp, cp, cp, ip, ip, ip]),
    "cntc_initialize": (None, [ip, ip, ip, ip, cp, ip]),
    "cntc_setglobalflags": (None, [ip, ip, ip]),
    "cntc_setflags": (None, [ip, ip, ip, ip, ip]),
    "cntc_setmetadata": (None, [ip, ip, ip, ip, dp]),
    "cntc_setsolverflags": (None, [ip, ip, ip, ip, ip, ip, dp]),
    "cntc_setmaterialparameters": (None, [ip, ip, ip, ip, dp]),
    "cntc_settimestep": (None, [ip, ip, dp]),
    "cntc_setreferencevelocity": (None, [ip, ip, dp]),
    "cntc_setrollingstepsize": (None, [ip, ip, dp, dp]),
    "cntc_setfrictionmethod": (None, [ip, ip, ip, ip, dp]),
    "cntc_sethertzcontact": (None, [ip, ip, ip, ip, dp]),
    "cntc_setpotcontact": (None, [ip, ip, ip, ip, dp]),
    "cntc_setpenetration": (None, [ip, ip, dp]),
    "cntc_setnormalforce": (None, [ip, ip, dp]),
    "cntc_setundeformeddistc": (None, [ip, ip, ip, ip, dp]),
    "cntc_setcreepages": (None, [ip, ip, dp, dp, dp]),
    "cntc_settangentialforces": (None, [ip, ip, dp, dp]),
    "subs_addblock": (None, [ip, ip, ip, ip, ip, ip, ip, dp, dp, dp]),
    "cntc_calculate": (None, [ip, ip, ip]),
    "subs_calculate": (None, [ip, ip, ip]),
    "cntc_getflags": (None, [ip, ip, ip, ip, ip]),
    "cntc_getnumelements": (None, [ip, ip, ip, ip]),
    "cntc_getgriddiscretization": (None, [ip, ip, dp, dp]),
    "cntc_getpotcontact": (None, [ip, ip, ip, dp]),
    "cntc_getpenetration": (None, [ip, ip, dp]),
    "cntc_getcreepages": (None, [ip, ip, dp, dp, dp]),
    "cntc_getcontactforces": (None, [ip, ip, dp, dp, dp, dp]),
    "cntc_getcontactpatchareas": (None, [ip, ip, dp, dp, dp]),
    "cntc_getelementdivision": (None, [ip, ip, ip, ip]),
    "cntc_getmaximumpressure": (None, [ip, ip, dp]),
    "cntc_getmaximumtraction": (None, [ip, ip, dp]),
    "cntc_getfielddata": (None, [ip, ip, ip, ip, dp]),
    "cntc_gettractions": (None, [ip, ip, ip, dp, dp, dp]),
    "cntc_getmicroslip": (None, [ip, ip, ip, dp, dp]),
    "cntc_getdisplacements": (None, [ip, ip, ip, dp, dp, dp]),
    "cntc_getcalculationtime": (None, [ip, ip, dp, dp]),
    "subs_getblocksize": (None, [ip, ip, ip, ip, ip, ip]),
    "subs_getresults": (None, [ip, ip, ip, ip, ip, ip, dp]),
    "cntc_getparameters": (None, [ip, ip, ip, ip, dp]),
    "cntc_getreferencevelocity": (None, [ip, ip, dp]),
    "cntc_gethertzcontact": (None, [ip, ip, ip, dp]),
    "cntc_getmaximumtemperature": (None, [ip, ip, dp, dp]),
    "cntc_getsensitivities": (None, [ip, ip, ip, ip, dp]),
    "cntc_resetcalculationtime": (None, [ip, ip]),
    "cntc_setextrarigidslip": (None, [ip, ip, ip, dp, dp]),
    "cntc_settemperaturedata": (None, [ip, ip, ip, ip, dp]),
    "cntc_readinpfile": (None, [ip, ip, cp, ip, ip]),
    "cntc_setverticalforce": (None, [ip, dp]),
    "cntc_setprofileinputfname": (None, [ip, cp, ip, ip, ip, ip, dp]),
    "cntc_setprofileinputvalues": (None, [ip, ip, dp, ip, ip, ip, dp]),
    "cntc_settrackdimensions": (None, [ip, ip, ip, dp]),
    "cntc_setwheelsetdimensions": (None, [ip, ip, ip, dp]),
    "cntc_setwheelsetposition": (None, [ip, ip, ip, dp]),
    "cntc_setwheelsetvelocity": (None, [ip, ip, ip, dp]),
    "cntc_setwheelsetflexibility": (None, [ip, ip, ip, dp]),
    "cntc_getprofilevalues": (None, [ip, ip, ip, ip, ip, dp, ip, dp]),
    "cntc_getprofilevalues_new": (None, [ip, ip, ip, ip, ip, dp, ip, dp]),
    "cntc_getwheelsetposition": (None, [ip, ip, dp]),
    "cntc_getwheelsetvelocity": (None, [ip, ip, dp]),
    "cntc_getnumcontactpatches": (None, [ip, ip]),
    "cntc_getcontactlocation": (None, [ip, ip, ip, dp]),
    "cntc_getglobalforces": (None, [ip, ip, ip, dp]),
    "cntc_finalize": (None, [ip]),
    "cntc_finalizelast": (None, []),
    "cntc_calculate_batch": (None, [ip, ip, ip, ip]),
    "cb200_get_iterations": (I, [I, I, ip, I, ip]),
    "cb200_get_outer_history": (I, [I, I, I, dp, dp]),
    "cb200_set_state": (I, [I, I, I, ip, dp]),
    "cb200_get_soutpt": (I, [I, I, I, dp]),
    "cb200_get_deformed_distance": (I, [I, I, I, dp]),
    "cb200_set_devices": (I, [I, ip]),
    "cb200_last_error": (C.c_char_p, []),
    "cb200_num_launches": (L, []),
    "cb200_num_sms": (I, []),
    "cb200_steady_prof": (I, [C.POINTER(C.c_ulonglong), I]),
    "cb200_gd_prof": (I, [C.POINTER(C.c_ulonglong), I]),
    "cb200_batch_timing": (I, [C.POINTER(C.c_double)]),
    "cb200_batch_timing_output": (I, [C.POINTER(C.c_double)]),
    "cb200_conv_prof": (I, [C.POINTER(C.c_ulonglong), I]),
    "cb200_work_counters": (I, [C.POINTER(C.c_ulonglong), I]),
    "cb200_solver_prof": (I, [C.POINTER(C.c_ulonglong), I]),
    "cb200_opt_fft_size": (I, [I]),
    "cb200_coefset_create": (I, [I, I, D, D, D, D, D, D, I, D, D]),
    "cb200_coefset_get_block": (I, [I, I, I, I, dp]),
    "cb200_coefset_plan": (I, [I, ip]),
    "cb200_vecaijpj": (I, [I, I, I, I, I, I, dp, ip, dp]),
    "cb200_aijpj": (I, [I, I, I, I, I, ip, dp, ip, dp]),
    "cb200_vecaijpj_dev": (I, [I, I, I, I, I, I, V, V, V, V]),
    "cb200_snorm_batch_dev": (I, [I, I, I, I, I, D, V, V, V, V, V, V]),
    "cb200_eldiv0": (I, [I, I, D, D, D, D, D, D, I, dp, I, D, D, dp, ip, dp]),
    "cb200_snorm_batch": (I, [I, I, I, I, I, D, dp, ip, dp, dp, dp]),
    "cb200_snorm_kernel_ms": (D, []),
    "cb200_fp64_peak_tflops": (D, [I]),
    "cb200_subsurf_batch_dev": (I, [I, I, I, dp, D, D, D, D, V, V, V]),
    "cb200_subsurf_batch": (I, [I, I, I, dp, D, D, D, D, dp, dp]),
    "cb200_subsurf_points": (I, [I, I, D, D, D, D, D, D, D, D, dp, I, dp, dp]),
    "cb200_snorm_workspace_bytes": (L, [I, I]),
}


def library_path():
    return _SO


def load_library():
    """Load the CUDA library; raises LibraryMissing (never falls back to a CPU path)."""
    global _dll
    if _dll is not None:
        return _dll
    if not os.path.exists(_SO):
        raise LibraryMissing("%s not found: run `python -m contact_b200.build` (nvcc, sm_100a). "
                             "contact_b200 has no CPU fallback." % _SO)
    dll = C.CDLL(_SO)
    for name, (res, args) in PROTOTYPES.items():
        f = getattr(dll, name)          # AttributeError if a declared symbol is not exported
        f.restype = res
        f.argtypes = args
    _dll = dll
    return dll


def last_error():
    return load_library().cb200_last_error().decode()
