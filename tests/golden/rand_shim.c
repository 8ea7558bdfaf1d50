/* Interposes libc rand() with a constant (LD_PRELOAD) while tests/golden/make_golden.py records what the REFERENCE's
 * sep_force_dpd computes: sep_rand() = rand()/(RAND_MAX+1.0) (reference include/sepmisc.h:60) becomes exactly 0.75 for
 * every pair, which SEPGPU_DPD_SEED_FIXED reproduces on the device and in the oracle.  Test infrastructure only. */
int rand(void) { return 1610612736; }   /* 3 * 2^29 */
