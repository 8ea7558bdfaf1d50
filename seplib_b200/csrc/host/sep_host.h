/* sep_host.h -- internals of the C99 host layer: the binding between a caller-owned seppart array and
 * its device context, and the coherence bookkeeping described at the top of include/sep.h. */
#ifndef SEP_HOST_H
#define SEP_HOST_H

#include "sep.h"
#include "sepgpu.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

/* bits for host_dirty (host newer than device) and dev_dirty (device newer than host) */
#define SEPB_X      (1u << 0)
#define SEPB_V      (1u << 1)
#define SEPB_F      (1u << 2)
#define SEPB_M      (1u << 3)
#define SEPB_Z      (1u << 4)
#define SEPB_TYPE   (1u << 5)
#define SEPB_MOL    (1u << 6)
#define SEPB_XN     (1u << 7)
#define SEPB_CN     (1u << 8)
#define SEPB_CR     (1u << 9)
#define SEPB_PV     (1u << 10)
#define SEPB_PA     (1u << 11)
#define SEPB_A      (1u << 12)
#define SEPB_EXCL   (1u << 13)
#define SEPB_TOPO   (1u << 14)
#define SEPB_X0     (1u << 15)
#define SEPB_ALL_STATE (SEPB_X | SEPB_V | SEPB_F | SEPB_M | SEPB_Z | SEPB_TYPE | SEPB_MOL | SEPB_XN | SEPB_CN | SEPB_CR)

typedef struct sep_binding {
    seppart *atoms;
    size_t npart;
    sepgpu_ctx *gpu;            /* created on the first hot call */
    sepmolinfo *molptr;         /* secondary key: survives by-value copies of sepsys */
    unsigned host_dirty;
    unsigned dev_dirty;
    int uploaded_once;
    int dpd_state_on_device;
    int fij_on;                  /* device molecule-molecule force table enabled */
    int x0_on_device;            /* tether positions (seppart.x0) are mirrored on the device */
    double *noise; size_t noise_cap;   /* per-step Gaussian numbers for sep_fp / sep_langevinGJF */
    double *alpha_ptr[4];       /* caller-owned thermostat multipliers mapped to device slots */
    double alpha_seen[4];
    unsigned long long dpd_calls;
    sepret *last_ret;
    double *blengths_host, *angles_host, *dihedrals_host;
    /* sep_force_pairs with a pair function of the caller's own: its table, sampled once per (function, cutoff) */
    struct sep_pairtab { double (*fun)(double, char); double cf, r2_lo; int n; double *fu; } pairtab[4];
    unsigned pairtab_next;
    /* SEP_SYNC=auto: atoms[] is an mmap'ed region whose protection tracks which side is newer */
    int managed;                 /* 1: allocated by sep_init in auto mode */
    void *map_base; size_t map_bytes;
    int prot;                    /* current protection of the region (PROT_NONE / PROT_READ / PROT_READ|PROT_WRITE) */
    int dd;                      /* SEP_NGPU > 1: this process drives one slab of the box (sep_dd.c) */
    int dd_rows_valid;
    struct sep_binding *next;
} sep_binding;

sep_binding *sepb_find(const seppart *atoms);
sep_binding *sepb_find_mol(const sepmolinfo *molptr);
sep_binding *sepb_first(void);
sep_binding *sepb_register(seppart *atoms, size_t npart);
void sepb_unregister(seppart *atoms);

/* make the device ready for a hot call: create the context, upload what the host changed */
sep_binding *sepb_prepare(seppart *atoms, sepsys *sys);
void sepb_fill_sys(const sepsys *sys, sepgpu_sys *out);
/* device -> host for the given field bits (only those currently newer on the device) */
void sepb_download(sep_binding *b, unsigned fields);
/* sepret / sepsys scalars device -> host */
void sepb_pull_scalars(sep_binding *b, sepsys *sys, sepret *ret, sepgpu_scalars *out);
void sepb_mark_host_dirty(seppart *atoms, unsigned fields);
void sepb_after_force(sep_binding *b, sepsys *sys, sepret *ret);
int sep_sync_mode(void);
/* atoms[] refreshed eagerly after integrator calls (step / full, or auto on an array the library cannot protect) */
int sepb_eager(const sep_binding *b);
/* the device copy of `bits` is newer than atoms[] from now on (auto mode: protects the array) */
void sepb_dev_newer(sep_binding *b, unsigned bits);
void sepb_check(int rc, const char *where);
unsigned long long sep_dpd_seed(void);

/* SEP_NGPU (sep_dd.c) */
int sepdd_world(void);
int sepdd_rank(void);
void sepdd_allow(void);                    /* the next sepb_prepare belongs to a call that runs decomposed */
void sepdd_guard(const sep_binding *b);
void sepdd_start(sep_binding *b, sepsys *sys);
void sepdd_before_upload(sep_binding *b);
void sepdd_download(sep_binding *b, int nf, const int *fl, const size_t *offs);
void sepdd_mark_failed(void);
FILE *sepdd_fopen(const char *path, const char *mode);
#define fopen(path, mode) sepdd_fopen(path, mode)     /* copies 1..N-1 of a SEP_NGPU run write nothing */

#endif
