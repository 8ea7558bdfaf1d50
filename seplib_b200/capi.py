"""ctypes bindings for libsep.so: the sepgpu_* device C ABI (include/sepgpu.h) and the seplib sep_* API
(include/sep.h).  Used by tests/ and bench.py; this is plumbing, not a compute path.  Everything here
fails loudly (RuntimeError) when the library or a CUDA device is missing -- there is no CPU fallback.

The Structure definitions mirror the reference's include/sepstrct.h:23-204 field for field, so the same
classes drive both our libsep.so and the compiled reference (oracle/_ref/libsep_ref.so) in the tests.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEPLIB_SO") or os.path.join(HERE, "libsep.so")      # SEPLIB_SO: A/B builds of the same library (scripts/)

SEP_BOND, SEP_ANGLE, SEP_DIHED = 10, 10, 20
SEP_ALL, SEP_EXCL_BONDED, SEP_EXCL_SAME_MOL = 1, 2, 3
SEP_BRUTE, SEP_NEIGHBLIST, SEP_LLIST_NEIGHBLIST = 0, 1, 2
SEP_LJCF2 = 0.016316891136

# sepgpu field ids (include/sepgpu.h)
(F_X, F_V, F_F, F_M, F_Z, F_TYPE, F_MOLINDEX, F_XN, F_CROSS_NEIGHB, F_CROSSINGS, F_PV, F_PA, F_A,
 F_BOND, F_ANGLE, F_DIHED, F_GID, F_X0) = range(18)

d3 = C.c_double * 3
i3 = C.c_int * 3
d33 = (C.c_double * 3) * 3


class SepPart(C.Structure):
    _fields_ = [
        ("x", d3), ("v", d3), ("f", d3), ("a", d3), ("m", C.c_double), ("type", C.c_char),
        ("z", C.c_double), ("neighb", C.POINTER(C.c_int)), ("cross_neighb", i3), ("crossings", i3),
        ("molindex", C.c_int), ("bond", C.c_int * SEP_BOND), ("angle", C.c_int * SEP_ANGLE),
        ("dihed", C.c_int * SEP_DIHED), ("sigma", C.c_double), ("collid", C.POINTER(C.c_int)),
        ("colltime", C.POINTER(C.c_double)), ("ldiff", C.c_double), ("xtrue", d3), ("x0", d3),
        ("xn", d3), ("xp", d3), ("px", d3), ("pv", d3), ("pa", d3), ("randn", d3), ("prevf", d3),
    ]


class SepMolInfo(C.Structure):
    _fields_ = [
        ("num_mols", C.c_uint), ("max_nuau", C.c_uint),
        ("flag_bonds", C.c_int), ("flag_angles", C.c_int), ("flag_dihedrals", C.c_int),
        ("num_bonds", C.c_uint), ("blist", C.POINTER(C.c_uint)), ("num_btypes", C.c_uint),
        ("num_angles", C.c_uint), ("alist", C.POINTER(C.c_uint)), ("num_atypes", C.c_uint),
        ("num_dihedrals", C.c_uint), ("dlist", C.POINTER(C.c_uint)), ("num_dtypes", C.c_uint),
        ("blengths", C.POINTER(C.c_double)), ("angles", C.POINTER(C.c_double)),
        ("dihedrals", C.POINTER(C.c_double)),
        ("flag_Fij", C.c_uint), ("Fij", C.c_void_p), ("Fiajb", C.c_void_p),
    ]


class SepSys(C.Structure):
    _fields_ = [
        ("npart", C.c_long), ("length", d3), ("volume", C.c_double),
        ("intgr_type", C.c_int), ("dt", C.c_double), ("tnow", C.c_double), ("ndof", C.c_uint),
        ("max_dist2", C.c_double),
        ("cf", C.c_double), ("lsubbox", d3), ("nsubbox", i3), ("skin", C.c_double),
        ("neighb_update", C.c_uint), ("neighb_flag", C.c_uint), ("nupdate_neighb", C.c_uint),
        ("omp_flag", C.c_bool), ("nthreads", C.c_uint), ("fun_cstate", C.c_int),
        ("molptr", C.POINTER(SepMolInfo)),
    ]


class SepRet(C.Structure):
    _fields_ = [
        ("etot", C.c_double), ("ekin", C.c_double), ("epot", C.c_double), ("ecoul", C.c_double),
        ("sumv2", C.c_double),
        ("P", d33), ("kin_P", d33), ("pot_P", d33), ("p", C.c_double),
        ("P_mol", d33), ("kin_P_mol", d33), ("pot_P_mol", d33), ("p_mol", C.c_double),
        ("pot_P_conservative", d33), ("pot_P_random", d33), ("pot_P_dissipative", d33),
        ("pot_P_bond", d33),
        ("pot_T_mol", d33), ("kin_T_mol", d33), ("T_mol", d33), ("t_mol", C.c_double),
    ]


assert C.sizeof(SepPart) == 568, C.sizeof(SepPart)
assert C.sizeof(SepRet) == 1000, C.sizeof(SepRet)


class GpuSys(C.Structure):
    _fields_ = [("length", d3), ("lsubbox", d3), ("nsubbox", i3), ("cf", C.c_double),
                ("skin", C.c_double), ("dt", C.c_double), ("neighb_update", C.c_int)]


class GpuLJ(C.Structure):
    _fields_ = [("cf", C.c_double), ("eps", C.c_double), ("sigma", C.c_double), ("aw", C.c_double),
                ("shift", C.c_double)]


class GpuScalars(C.Structure):
    _fields_ = [("epot", C.c_double), ("ecoul", C.c_double), ("ekin", C.c_double),
                ("pot_P", C.c_double * 9), ("kin_P", C.c_double * 9), ("pot_P_bond", C.c_double * 9),
                ("max_dist2", C.c_double), ("sum_mv2", C.c_double), ("alpha", C.c_double * 4),
                ("neighb_flag", C.c_int), ("nbuild", C.c_int), ("error", C.c_int),
                ("max_neighb", C.c_int), ("npairs_listed", C.c_longlong)]


PAIRFUN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_char)

# numpy view of a seppart array (host AoS, 568-byte records) for bulk reads/writes in tests and bench
ATOM_DTYPE = np.dtype({
    "names": ["x", "v", "f", "a", "m", "type", "z", "cross_neighb", "crossings", "molindex", "bond",
              "angle", "dihed", "xn", "pv", "pa"],
    "formats": [(np.float64, 3), (np.float64, 3), (np.float64, 3), (np.float64, 3), np.float64, np.uint8,
                np.float64, (np.int32, 3), (np.int32, 3), np.int32, (np.int32, 10), (np.int32, 10),
                (np.int32, 20), (np.float64, 3), (np.float64, 3), (np.float64, 3)],
    "offsets": [0, 24, 48, 72, 96, 104, 112, 128, 140, 152, 156, 196, 236, 400, 472, 496],
    "itemsize": 568,
})


def atoms_view(ptr, n):
    """Structured numpy view (no copy) of the seppart array returned by sep_init / sep_init_xyz."""
    addr = C.addressof(ptr.contents)
    buf = (C.c_char * (568 * n)).from_address(addr)
    return np.frombuffer(buf, dtype=ATOM_DTYPE, count=n)

# every symbol include/sepgpu.h declares (tests check that the library exports all of them)
SEPGPU_SYMBOLS = [
    "sepgpu_create", "sepgpu_destroy", "sepgpu_last_error", "sepgpu_device_count", "sepgpu_put",
    "sepgpu_get", "sepgpu_put_fields", "sepgpu_get_fields", "sepgpu_set_topology", "sepgpu_get_bonded_values", "sepgpu_reset_ret",
    "sepgpu_reset_force", "sepgpu_neighb_build", "sepgpu_force_lj", "sepgpu_force_table", "sepgpu_set_host_rows", "sepgpu_md_lj_nvt", "sepgpu_dd_set_charges", "sepgpu_coulomb_sf",
    "sepgpu_feed_vacf", "sepgpu_feed_msd", "sepgpu_feed_profile", "sepgpu_feed_fourier", "sepgpu_feed_radial",
    "sepgpu_force_dpd", "sepgpu_stretch_harmonic", "sepgpu_angle_harmonic", "sepgpu_angle_cossq",
    "sepgpu_torsion_ryckaert", "sepgpu_bonded_side", "sepgpu_nosehoover", "sepgpu_nosehoover_type", "sepgpu_set_alpha",
    "sepgpu_leapfrog", "sepgpu_verlet_dpd", "sepgpu_reset_momentum", "sepgpu_scale_positions",
    "sepgpu_read_scalars", "sepgpu_sync", "sepgpu_get_pairs", "sepgpu_request_rebuild",
    "sepgpu_set_option", "sepgpu_get_option", "sepgpu_timer_start", "sepgpu_timer_stop", "sepgpu_kernel_time",
    "sepgpu_peak_fp64", "sepgpu_peak_copy", "sepgpu_flush_l2",
    "sepgpu_fij_enable", "sepgpu_fij_reset", "sepgpu_fij_get", "sepgpu_scale_box", "sepgpu_relax_temp", "sepgpu_force_x0", "sepgpu_fp", "sepgpu_langevin_gjf",
    "sepgpu_dd_unique_id", "sepgpu_dd_init", "sepgpu_dd_set_owned", "sepgpu_dd_layers",
]

# the sep_* symbols include/sep.h declares
SEP_SYMBOLS = [
    "sep_init", "sep_close", "sep_init_xyz", "sep_sys_setup", "sep_free_sys", "sep_set_lattice",
    "sep_set_vel", "sep_set_vel_seed", "sep_set_vel_type", "sep_force_pairs", "sep_force_lj",
    "sep_force_dpd", "sep_neighb", "sep_neighb_nonbonded", "sep_neighb_excl_same_mol",
    "sep_bond_share", "sep_angle_share", "sep_dihed_share", "sep_bonded", "sep_coulomb_sf",
    "sep_periodic", "sep_leapfrog", "sep_nosehoover", "_sep_nosehoover_type", "sep_verlet_dpd",
    "sep_read_topology_file", "sep_free_bonds", "sep_free_angles", "sep_free_dihedrals",
    "sep_init_mol", "sep_free_mol", "sep_stretch_harmonic", "sep_angle_harmonic", "sep_angle_cossq",
    "sep_torsion_Ryckaert", "sep_mol_cm", "sep_mol_velcm", "sep_eval_mol_pressure_tensor",
    "sep_average_bondlengths", "sep_reset_retval", "sep_get_pressure", "sep_get_temperature",
    "sep_pressure_tensor", "sep_mol_pressure_tensor", "sep_error", "sep_warning", "sep_lj",
    "sep_lj_shift", "sep_wca", "sep_pairs_retabulate", "sep_reset_force", "sep_reset_force_mol", "sep_nsubbox",
    "sep_box_length", "sep_count_type", "sep_set_x0", "sep_set_xn", "sep_save_xyz", "sep_eval_mom",
    "sep_eval_mom_type", "sep_compress_box", "sep_set_charge", "sep_set_mass", "sep_set_type",
    "sep_set_omp", "sep_set_skin", "sep_set_ndof", "sep_reset_momentum", "sep_dist_ij",
    "sep_eval_xtrue", "sep_vector", "sep_vector_int", "sep_matrix", "sep_free_matrix", "sep_matrix_set", "sep_omp_bond", "sep_omp_angle", "sep_omp_torsion",
    "sep_tensor_float", "sep_free_tensor_float", "sep_dot", "sep_vector_set", "sep_init_sampler",
    "sep_add_sampler", "sep_add_mol_sampler", "sep_sample", "sep_close_sampler", "sep_gpu_set_sync",
    "sep_gpu_sync", "sep_gpu_invalidate", "sep_gpu_sync_scalars", "sep_gpu_export_neighb",
    "sep_gpu_handle", "sep_gpu_set_dpd_seed",
    "sep_compress_box_dir", "sep_compress_box_dir_length", "sep_berendsen", "sep_berendsen_iso", "sep_relax_temp",
    "sep_spring_x0", "sep_force_x0", "sep_mol_eval_xtrue", "sep_mol_spin", "sep_mol_dipoles",
    "sep_randn", "sep_fp", "sep_langevinGJF", "sep_set_ldiff",
]


def declare_sep_api(lib):
    """argtypes/restypes of the sep_* functions the tests call; valid for libsep.so and libsep_ref.so."""
    P, S, R = C.POINTER(SepPart), C.POINTER(SepSys), C.POINTER(SepRet)
    lib.sep_init.restype = P
    lib.sep_init.argtypes = [C.c_size_t, C.c_size_t]
    lib.sep_close.argtypes = [P, C.c_size_t]
    lib.sep_init_xyz.restype = P
    lib.sep_init_xyz.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_char]
    lib.sep_sys_setup.restype = SepSys
    lib.sep_sys_setup.argtypes = [C.c_double] * 5 + [C.c_size_t, C.c_size_t]
    lib.sep_free_sys.argtypes = [S]
    lib.sep_set_lattice.argtypes = [P, SepSys]
    lib.sep_set_vel_seed.argtypes = [P, C.c_double, C.c_uint, SepSys]
    lib.sep_reset_retval.argtypes = [R]
    lib.sep_reset_force.argtypes = [P, S]
    lib.sep_reset_force_mol.argtypes = [S]
    lib.sep_force_pairs.restype = C.c_int
    lib.sep_force_pairs.argtypes = [P, C.c_char_p, C.c_double, C.c_void_p, S, R, C.c_uint]
    lib.sep_force_lj.argtypes = [P, C.c_char_p, C.POINTER(C.c_double), S, R, C.c_uint]
    lib.sep_force_dpd.argtypes = [P, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double, S, R, C.c_uint]
    lib.sep_coulomb_sf.argtypes = [P, C.c_double, S, R, C.c_uint]
    lib.sep_nosehoover.argtypes = [P, C.c_double, C.POINTER(C.c_double), C.c_double, S]
    lib._sep_nosehoover_type.argtypes = [P, C.c_char, C.c_double, C.POINTER(C.c_double), C.c_double, S]
    lib.sep_leapfrog.argtypes = [P, S, R]
    lib.sep_verlet_dpd.argtypes = [P, C.c_double, C.c_int, S, R]
    lib.sep_read_topology_file.argtypes = [P, C.c_char_p, S, C.c_char]
    lib.sep_init_mol.restype = C.c_void_p
    lib.sep_init_mol.argtypes = [P, S]
    lib.sep_stretch_harmonic.argtypes = [P, C.c_int, C.c_double, C.c_double, S, R]
    lib.sep_angle_harmonic.argtypes = [P, C.c_int, C.c_double, C.c_double, S, R]
    lib.sep_angle_cossq.argtypes = [P, C.c_int, C.c_double, C.c_double, S, R]
    lib.sep_torsion_Ryckaert.argtypes = [P, C.c_int, C.POINTER(C.c_double), S, R]
    lib.sep_pressure_tensor.argtypes = [R, S]
    lib.sep_eval_mom.restype = C.c_double
    lib.sep_eval_mom.argtypes = [P, C.c_int]
    lib.sep_compress_box.argtypes = [P, C.c_double, C.c_double, S]
    lib.sep_compress_box_dir.argtypes = [P, C.c_double, C.c_double, C.c_int, S]
    lib.sep_compress_box_dir_length.argtypes = [P, C.c_double, C.c_double, C.c_int, S]
    lib.sep_berendsen.argtypes = [P, C.c_double, C.c_double, R, S]
    lib.sep_berendsen_iso.argtypes = [P, C.c_double, C.c_double, R, S]
    lib.sep_relax_temp.argtypes = [P, C.c_char, C.c_double, C.c_double, S]
    lib.sep_force_x0.argtypes = [P, C.c_char, C.c_void_p, S]
    lib.sep_set_x0.argtypes = [P, C.c_int]
    lib.sep_fp.argtypes = [P, C.c_double, S, R]
    lib.sep_langevinGJF.argtypes = [P, C.c_double, C.c_double, S, R]
    lib.sep_randn.restype = C.c_double
    lib.sep_reset_momentum.argtypes = [P, C.c_char, S]
    lib.sep_set_skin.argtypes = [S, C.c_double]
    lib.sep_set_omp.argtypes = [C.c_uint, S]
    lib.sep_neighb.argtypes = [P, S]
    lib.sep_neighb_nonbonded.argtypes = [P, S]
    lib.sep_neighb_excl_same_mol.argtypes = [P, S]
    return lib


_lib = None


def load():
    """Load libsep.so (built in-tree by seplib_b200/Makefile).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(seplib-b200 has no fallback implementation)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    ctx = C.c_void_p
    lib.sepgpu_last_error.restype = C.c_char_p
    lib.sepgpu_create.argtypes = [C.POINTER(ctx), C.c_size_t, C.c_int]
    lib.sepgpu_destroy.argtypes = [ctx]
    lib.sepgpu_destroy.restype = None
    lib.sepgpu_put.argtypes = [ctx, C.c_int, C.c_void_p, C.c_size_t]
    lib.sepgpu_get.argtypes = [ctx, C.c_int, C.c_void_p, C.c_size_t]
    lib.sepgpu_set_topology.argtypes = [ctx, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint]
    lib.sepgpu_get_bonded_values.argtypes = [ctx, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sepgpu_reset_ret.argtypes = [ctx]
    lib.sepgpu_reset_force.argtypes = [ctx]
    lib.sepgpu_neighb_build.argtypes = [ctx, C.POINTER(GpuSys), C.c_uint]
    lib.sepgpu_force_lj.argtypes = [ctx, C.POINTER(GpuSys), C.c_char_p, C.POINTER(GpuLJ), C.c_uint, C.c_int]
    lib.sepgpu_md_lj_nvt.argtypes = [ctx, C.POINTER(GpuSys), C.c_char_p, C.POINTER(GpuLJ), C.c_uint, C.c_double, C.c_int, C.c_double, C.c_int]
    lib.sepgpu_force_table.argtypes = [ctx, C.POINTER(GpuSys), C.c_char_p, C.c_double, C.POINTER(C.c_double), C.c_int, C.c_double,
                                       C.c_uint, C.c_int]
    lib.sepgpu_coulomb_sf.argtypes = [ctx, C.POINTER(GpuSys), C.c_double, C.c_uint]
    lib.sepgpu_force_dpd.argtypes = [ctx, C.POINTER(GpuSys), C.c_char_p, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_uint, C.c_ulonglong, C.c_ulonglong]
    lib.sepgpu_stretch_harmonic.argtypes = [ctx, C.POINTER(GpuSys), C.c_int, C.c_double, C.c_double]
    lib.sepgpu_angle_harmonic.argtypes = [ctx, C.POINTER(GpuSys), C.c_int, C.c_double, C.c_double]
    lib.sepgpu_angle_cossq.argtypes = [ctx, C.POINTER(GpuSys), C.c_int, C.c_double, C.c_double]
    lib.sepgpu_torsion_ryckaert.argtypes = [ctx, C.POINTER(GpuSys), C.c_int, C.POINTER(C.c_double)]
    lib.sepgpu_bonded_side.argtypes = [ctx, C.POINTER(GpuSys), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.sepgpu_nosehoover.argtypes = [ctx, C.POINTER(GpuSys), C.c_double, C.c_int, C.c_double]
    lib.sepgpu_nosehoover_type.argtypes = [ctx, C.POINTER(GpuSys), C.c_char, C.c_double, C.POINTER(C.c_double), C.c_double]
    lib.sepgpu_set_alpha.argtypes = [ctx, C.c_int, C.c_double]
    lib.sepgpu_leapfrog.argtypes = [ctx, C.POINTER(GpuSys)]
    lib.sepgpu_verlet_dpd.argtypes = [ctx, C.POINTER(GpuSys), C.c_double, C.c_int]
    lib.sepgpu_reset_momentum.argtypes = [ctx, C.c_char]
    lib.sepgpu_scale_positions.argtypes = [ctx, C.c_double]
    lib.sepgpu_read_scalars.argtypes = [ctx, C.POINTER(GpuScalars)]
    lib.sepgpu_sync.argtypes = [ctx]
    lib.sepgpu_get_pairs.restype = C.c_longlong
    lib.sepgpu_get_pairs.argtypes = [ctx, C.c_void_p, C.c_longlong]
    lib.sepgpu_request_rebuild.argtypes = [ctx]
    lib.sepgpu_set_option.argtypes = [ctx, C.c_char_p, C.c_longlong]
    lib.sepgpu_get_option.argtypes = [ctx, C.c_char_p, C.POINTER(C.c_longlong)]
    lib.sepgpu_timer_start.argtypes = [ctx]
    lib.sepgpu_timer_stop.argtypes = [ctx, C.POINTER(C.c_float)]
    lib.sepgpu_kernel_time.argtypes = [ctx, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.sepgpu_peak_fp64.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.sepgpu_peak_copy.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.sepgpu_flush_l2.argtypes = [ctx]
    lib.sepgpu_scale_box.argtypes = [ctx, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.sepgpu_relax_temp.argtypes = [ctx, C.POINTER(GpuSys), C.c_char, C.c_double, C.c_double, C.POINTER(C.c_double)]
    lib.sepgpu_force_x0.argtypes = [ctx, C.POINTER(GpuSys), C.c_char, C.c_double]
    lib.sepgpu_fp.argtypes = [ctx, C.POINTER(GpuSys), C.c_double, C.c_void_p]
    lib.sepgpu_langevin_gjf.argtypes = [ctx, C.POINTER(GpuSys), C.c_double, C.c_double, C.c_void_p]
    lib.sepgpu_feed_vacf.argtypes = [ctx, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.sepgpu_feed_msd.argtypes = [ctx, C.c_int, C.c_char, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sepgpu_feed_profile.argtypes = [ctx, C.c_char, C.c_double, C.c_int, C.c_void_p]
    lib.sepgpu_feed_fourier.argtypes = [ctx, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    lib.sepgpu_feed_radial.argtypes = [ctx, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
    lib.sepgpu_fij_enable.argtypes = [ctx, C.c_int]
    lib.sepgpu_fij_reset.argtypes = [ctx]
    lib.sepgpu_fij_get.argtypes = [ctx, C.c_void_p]
    lib.sepgpu_dd_unique_id.argtypes = [C.c_void_p]
    lib.sepgpu_dd_init.argtypes = [ctx, C.c_int, C.c_int, C.c_void_p, C.POINTER(GpuSys), C.c_longlong]
    lib.sepgpu_dd_set_owned.argtypes = [ctx, C.c_int]
    lib.sepgpu_dd_layers.argtypes = [ctx, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    declare_sep_api(lib)
    lib.sep_gpu_set_sync.argtypes = [C.c_int]
    lib.sep_gpu_sync.argtypes = [C.POINTER(SepPart)]
    lib.sep_gpu_handle.restype = C.c_void_p
    lib.sep_gpu_handle.argtypes = [C.POINTER(SepPart)]
    lib.sep_gpu_export_neighb.restype = C.c_long
    lib.sep_gpu_export_neighb.argtypes = [C.POINTER(SepPart), C.POINTER(SepSys), C.c_void_p, C.c_long]
    _lib = lib
    return lib


class GpuError(RuntimeError):
    pass


def _ck(lib, rc, what):
    if rc != 0:
        raise GpuError(f"{what} failed ({rc}): {lib.sepgpu_last_error().decode()}")


def make_sys(length, cf, dt, neighb_update=SEP_LLIST_NEIGHBLIST, skin=0.25, grid_skin=0.25):
    """sepgpu_sys with the cell grid sep_sys_setup would derive (reference source/sepinit.c:257-276)."""
    s = GpuSys()
    for k in range(3):
        s.length[k] = float(length[k])
        n = int(float(length[k]) / (cf + grid_skin))
        s.nsubbox[k] = n
        s.lsubbox[k] = float(length[k]) / n if n > 0 else 0.0
    s.cf, s.skin, s.dt, s.neighb_update = cf, skin, dt, neighb_update
    return s


class System:
    """Thin object wrapper over a sepgpu context, SoA numpy arrays in, numpy arrays out."""

    def __init__(self, npart, device=-1):
        self.lib = load()
        self.n = int(npart)
        self.ctx = C.c_void_p()
        _ck(self.lib, self.lib.sepgpu_create(C.byref(self.ctx), self.n, device), "sepgpu_create")

    def close(self):
        if self.ctx:
            self.lib.sepgpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    _dtypes = {F_TYPE: (np.uint8, 1), F_MOLINDEX: (np.int32, 1), F_GID: (np.int32, 1), F_CROSS_NEIGHB: (np.int32, 3),
               F_CROSSINGS: (np.int32, 3), F_BOND: (np.int32, 10), F_ANGLE: (np.int32, 10),
               F_DIHED: (np.int32, 20), F_M: (np.float64, 1), F_Z: (np.float64, 1)}

    def put(self, field, arr):
        dt, w = self._dtypes.get(field, (np.float64, 3))
        a = np.ascontiguousarray(arr, dtype=dt).reshape(self.n, w) if w > 1 else np.ascontiguousarray(arr, dtype=dt).reshape(self.n)
        assert len(a) == self.n
        _ck(self.lib, self.lib.sepgpu_put(self.ctx, field, a.ctypes.data, 0), f"sepgpu_put({field})")

    def get(self, field):
        dt, w = self._dtypes.get(field, (np.float64, 3))
        a = np.empty((self.n, w) if w > 1 else (self.n,), dtype=dt)
        _ck(self.lib, self.lib.sepgpu_get(self.ctx, field, a.ctypes.data, 0), f"sepgpu_get({field})")
        return a

    # ---- slab domain decomposition (one process per GPU) -------------------------------------------
    def dd_init(self, rank, world, id_bytes, gsys, n_global):
        buf = (C.c_char * 128).from_buffer_copy(id_bytes)
        _ck(self.lib, self.lib.sepgpu_dd_init(self.ctx, rank, world, buf, C.byref(gsys), n_global), "sepgpu_dd_init")
        self.ncap = self.n

    def dd_set_owned(self, n_own):
        _ck(self.lib, self.lib.sepgpu_dd_set_owned(self.ctx, n_own), "sepgpu_dd_set_owned")
        self.n = int(n_own)

    def dd_layers(self):
        z0, z1, no, nh = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _ck(self.lib, self.lib.sepgpu_dd_layers(self.ctx, C.byref(z0), C.byref(z1), C.byref(no), C.byref(nh)), "sepgpu_dd_layers")
        self.n = no.value
        return z0.value, z1.value, no.value, nh.value

    def call(self, name, *args):
        _ck(self.lib, getattr(self.lib, name)(self.ctx, *args), name)

    def scalars(self):
        s = GpuScalars()
        _ck(self.lib, self.lib.sepgpu_read_scalars(self.ctx, C.byref(s)), "sepgpu_read_scalars")
        return s

    def pairs(self, max_pairs=None):
        if max_pairs is None:
            max_pairs = max(1024, int(self.scalars().npairs_listed) // 2 + 16)
        buf = np.empty((max_pairs, 2), dtype=np.int32)
        n = self.lib.sepgpu_get_pairs(self.ctx, buf.ctypes.data, max_pairs)
        if n < 0:
            raise GpuError(f"sepgpu_get_pairs failed ({n}): {self.lib.sepgpu_last_error().decode()}")
        return buf[:n]

    def set_topology(self, blist, alist, dlist):
        b = np.ascontiguousarray(blist, dtype=np.uint32).reshape(-1, 3)
        a = np.ascontiguousarray(alist, dtype=np.uint32).reshape(-1, 4)
        d = np.ascontiguousarray(dlist, dtype=np.uint32).reshape(-1, 5)
        _ck(self.lib, self.lib.sepgpu_set_topology(self.ctx, b.ctypes.data, len(b), a.ctypes.data, len(a),
                                                   d.ctypes.data, len(d)), "sepgpu_set_topology")
        self._nb, self._na, self._nd = len(b), len(a), len(d)

    def bonded_values(self):
        bl = np.zeros(max(self._nb, 1)); an = np.zeros(max(self._na, 1)); di = np.zeros(max(self._nd, 1))
        _ck(self.lib, self.lib.sepgpu_get_bonded_values(self.ctx, bl.ctypes.data, an.ctypes.data, di.ctypes.data),
            "sepgpu_get_bonded_values")
        return bl[:self._nb], an[:self._na], di[:self._nd]


def dd_unique_id():
    """128-byte NCCL unique id (call on rank 0, broadcast to the others)."""
    lib = load()
    buf = (C.c_char * 128)()
    _ck(lib, lib.sepgpu_dd_unique_id(buf), "sepgpu_dd_unique_id")
    return bytes(buf)


def dd_slab_range(rank, world, nz):
    """Global cell layers [z0,z1) owned by `rank` -- the same split as sepgpu_dd_init."""
    return rank * nz // world, (rank + 1) * nz // world


def lj_param(cf, eps=1.0, sigma=1.0, aw=1.0, shift=None, kind=None):
    """kind in {None,'lj','lj_shift','wca','param'} -> sepgpu_ljparam with the reference's shift rules."""
    p = GpuLJ()
    p.cf, p.eps, p.sigma, p.aw = cf, eps, sigma, aw
    if shift is not None:
        p.shift = shift
    elif kind in (None, "lj"):
        p.shift = 0.0
    elif kind == "lj_shift":
        p.shift = -SEP_LJCF2
    elif kind == "wca":
        p.shift = -1.0
    elif kind == "param":
        p.shift = 4.0 * eps * ((sigma / cf) ** 12 - aw * (sigma / cf) ** 6)
    else:
        raise ValueError(kind)
    return p
