/* Shim for <cint.h> (libcint v6.1.3 is NOT vendored in the reference tree): just the declarations the
 * reference's own driver/digestion sources need — slot macros and an opaque CINTOpt.  Integrals enter those
 * sources only through function pointers (pyscf/lib/vhf/nr_direct.c:73), which oracle/ supplies. */
#ifndef ORACLE_SHIM_CINT_H
#define ORACLE_SHIM_CINT_H
#define CHARGE_OF 0
#define PTR_COORD 1
#define NUC_MOD_OF 2
#define PTR_ZETA 3
#define PTR_FRAC_CHARGE 4
#define ATM_SLOTS 6
#define ATOM_OF 0
#define ANG_OF 1
#define NPRIM_OF 2
#define NCTR_OF 3
#define KAPPA_OF 4
#define PTR_EXP 5
#define PTR_COEFF 6
#define BAS_SLOTS 8
#define PTR_EXPCUTOFF 0
#define PTR_COMMON_ORIG 1
#define PTR_RINV_ORIG 4
#define PTR_RINV_ZETA 7
#define PTR_RANGE_OMEGA 8
#define PTR_ENV_START 20
#define NGRIDS 11
#define PTR_GRIDS 12
#ifndef MIN
#define MIN(X, Y) ((X) < (Y) ? (X) : (Y))
#endif
#ifndef MAX
#define MAX(X, Y) ((X) > (Y) ? (X) : (Y))
#endif
#define atm(SLOT, I) atm[ATM_SLOTS * (I) + (SLOT)]
#define bas(SLOT, I) bas[BAS_SLOTS * (I) + (SLOT)]
#define FINT int
typedef struct CINTOpt_s CINTOpt;
#endif
