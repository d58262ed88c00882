/* integration/dwgsim_b200_binding.h -- the binding a maintainer of nh13/DWGSIM adds to src/ to run the read-pair loop of
 * dwgsim_core (src/dwgsim.c:636-1099) on libdwgsim_b200.so.  Compiled against the reference's own headers (mut.h,
 * regions_bed.h, dwgsim_opt.h) and include/dwgsim_gpu.h; oracle/patch_reference.py + `make -C oracle ref_gpu` build a
 * reference binary with it (oracle/_ref/dwgsim_ref_gpu), tests/test_gpu_integration.py runs it. */
#ifndef DWGSIM_B200_BINDING_H
#define DWGSIM_B200_BINDING_H
#include <stdint.h>
#include <stdio.h>
#include <zlib.h>
#include "contigs.h"
#include "mut.h"
#include "regions_bed.h"
#include "dwgsim_opt.h"

/* replaces `for (ii = 0; ii != n_pairs; ++ii, ++ctr) {...}` for one contig: queues the contig as dwgsim_core holds it after
 * mut_diref and runs its pairs; FASTQ bytes go to opt->fp_bwa1 / fp_bwa2 / fp_bfast (gzFile).  `l` is the length the
 * sampler draws in (the total region length with -x, src/dwgsim.c:553).  Exits like the reference on errors. */
void dwgsim_b200_contig(dwgsim_opt_t *opt, int contig_i, const char *name, const seq_t *seq, const mutseq_t *hap1,
                        const mutseq_t *hap2, int64_t n_pairs, const regions_bed_txt *regions_bed, int l);
/* before dwgsim_core returns */
void dwgsim_b200_close(void);
#endif
