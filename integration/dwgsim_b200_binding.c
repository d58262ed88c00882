/* integration/dwgsim_b200_binding.c -- see dwgsim_b200_binding.h.  This file is the C side of INTEGRATION.md. */
#include <stdlib.h>
#include <string.h>
#include "dwgsim_b200_binding.h"
#include "dwgsim_gpu.h"

static dwgsim_gpu_t *g_gpu = NULL;

/* gzFile writer: the sink receives FASTQ bytes in pair order, per file id */
static int gz_sink(void *user, int file_id, const char *buf, size_t n)
{
    dwgsim_opt_t *opt = (dwgsim_opt_t *)user;
    gzFile fp = file_id == DWGSIM_GPU_FILE_BWA1 ? opt->fp_bwa1 : (file_id == DWGSIM_GPU_FILE_BWA2 ? opt->fp_bwa2 : opt->fp_bfast);
    while (fp != NULL && n > 0) {                       /* gzwrite takes an unsigned length */
        const unsigned chunk = n > (1u << 30) ? (1u << 30) : (unsigned)n;
        if (gzwrite(fp, buf, chunk) != (int)chunk) return 1;
        buf += chunk; n -= chunk;
    }
    return 0;
}

static void gpu_open(const dwgsim_opt_t *opt)
{
    dwgsim_gpu_params_t p;
    int i, rc;
    memset(&p, 0, sizeof p);
    for (i = 0; i < 2; i++) { p.e_start[i] = opt->e[i].start; p.e_by[i] = opt->e[i].by; p.length[i] = opt->length[i]; }
    p.is_inner = opt->is_inner; p.dist = opt->dist; p.std_dev = opt->std_dev;
    p.mut_freq = opt->mut_freq; p.rand_read = opt->rand_read; p.max_n = opt->max_n;
    p.data_type = opt->data_type; p.strandedness = opt->strandedness; p.read_one_strand = opt->read_one_strand;
    p.flow_order = opt->flow_order; p.flow_order_len = opt->flow_order_len;      /* already codes 0..3 (src/dwgsim_opt.c:404-407) */
    p.seed = opt->seed;
    p.fixed_quality = opt->fixed_quality ? (unsigned char)opt->fixed_quality[0] : 0;
    p.quality_std = opt->quality_std; p.read_prefix = opt->read_prefix;
    p.reads_output_type = opt->reads_output_type; p.amplicons = opt->amplicons;
    if ((rc = dwgsim_gpu_create(&g_gpu, &p, 0)) != DWGSIM_GPU_OK) {
        fprintf(stderr, "\n[dwgsim_core] Error: %s\n", dwgsim_gpu_strerror(rc));
        exit(1);                                        /* the reference's convention, src/dwgsim.c:177-180 */
    }
}

void dwgsim_b200_contig(dwgsim_opt_t *opt, int contig_i, const char *name, const seq_t *seq, const mutseq_t *hap1,
                        const mutseq_t *hap2, int64_t n_pairs, const regions_bed_txt *regions_bed, int l)
{
    int rc;
    if (n_pairs <= 0) return;
    if (g_gpu == NULL) gpu_open(opt);
    rc = dwgsim_gpu_add_contig(g_gpu, contig_i, name, seq->s, seq->l, (const uint64_t *)hap1->s, (const uint64_t *)hap2->s,
                               hap1->ins, hap1->ins_l, hap2->ins, hap2->ins_l, n_pairs);
    if (rc == DWGSIM_GPU_OK && regions_bed != NULL) {
        /* the contig's merged regions (src/regions_bed.c:43-115: sorted by contig and start) and the sampler's `l` */
        uint32_t first = 0, n = 0, r;
        for (r = 0; r < regions_bed->n; r++) if (regions_bed->contig[r] == (uint32_t)contig_i) { if (!n) first = r; n++; }
        rc = dwgsim_gpu_set_regions(g_gpu, regions_bed->start + first, regions_bed->end + first, (int32_t)n, l);
    }
    if (rc == DWGSIM_GPU_OK) rc = dwgsim_gpu_run(g_gpu, gz_sink, opt, NULL);
    if (rc != DWGSIM_GPU_OK) {
        fprintf(stderr, "\r[dwgsim_core] %s: %s\n", dwgsim_gpu_strerror(rc), dwgsim_gpu_last_error(g_gpu));
        exit(1);
    }
}

void dwgsim_b200_close(void)
{
    if (g_gpu != NULL) dwgsim_gpu_destroy(g_gpu);
    g_gpu = NULL;
}
