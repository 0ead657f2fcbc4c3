/* TEST INFRASTRUCTURE ONLY: the <config.h> htslib's kfunc.c includes (htslib's configure is not run). */
#define HAVE_FSEEKO 1
#define HAVE_DRAND48 1
