/*
 * bam_min.c -- TEST INFRASTRUCTURE ONLY.
 * The three htslib record functions BamAlignment's constructors / destructor call (bam_io.h:57-92), so that the parity
 * harness can build BamAlignment objects in memory and hand them to the reference's realign() / convertAlignment() /
 * TrimAlignment().  htslib itself is not built; these follow htslib's documented semantics for an in-memory bam1_t.
 */
#include <stdlib.h>
#include <string.h>

#include "htslib/sam.h"

bam1_t* bam_init1(void) { return (bam1_t*)calloc(1, sizeof(bam1_t)); }

void bam_destroy1(bam1_t* b) {
  if (b == NULL) return;
  free(b->data);
  free(b);
}

bam1_t* bam_copy1(bam1_t* dst, const bam1_t* src) {
  uint8_t* data = dst->data;
  if (dst->m_data < src->l_data) {
    data = (uint8_t*)realloc(data, src->l_data > 0 ? src->l_data : 1);
    dst->m_data = src->l_data;
  }
  if (src->l_data > 0) memcpy(data, src->data, src->l_data);
  {
    uint32_t m = dst->m_data;
    *dst = *src;
    dst->m_data = m;
    dst->data = data;
  }
  return dst;
}
