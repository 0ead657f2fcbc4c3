/* stands in for the file htslib's Makefile generates (version.sh: 1.9) */
#define HTS_VERSION "1.9"
