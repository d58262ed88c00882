// gz_device.cuh -- device side of the gzip writer (format and tables: gz_host.h).
//
//   gz_histogram_kernel   byte histogram of a raw FASTQ stream (once per handle: the per-stream Huffman code is
//                         fitted to the first batch; FASTQ symbol statistics are stationary)
//   gz_compress_kernel    one CTA per 64 KiB member, one thread per 256-byte chunk: code lengths -> bit offsets
//                         (block scan) -> LSB-first bit packing straight into the member's slot; CRC-32 by
//                         slicing-by-4 per chunk and a GF(2) tree combine; gzip header / trailer
//   gz_compact_kernel     members (variable size) -> one contiguous stream at the scanned offsets
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gz_host.h"

namespace dwg {

constexpr int kGzThreads = kGzMemberRaw / kGzChunk;        // 256

struct GzDeviceTables {                                    // per stream, in HBM
    const uint32_t *code;                                  // [257] (reversed code << 4) | length
    const uint8_t *prefix;                                 // gzip header + block header bits, padded to whole words
    uint32_t prefix_bits;
};

__global__ void __launch_bounds__(256)
gz_histogram_kernel(const uint8_t *__restrict__ raw, unsigned long long n, unsigned long long *__restrict__ hist /* [256] */)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * 16;
    for (unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
        if (base + 16 <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(raw + base);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int b = 0; b < 16; ++b) atomicAdd(&h[(w[b >> 2] >> (8 * (b & 3))) & 0xFFu], 1u);
        } else for (unsigned long long x = base; x < n; ++x) atomicAdd(&h[raw[x]], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// a(x) * b(x) mod P in the reflected representation (zlib's multmodp)
__device__ __forceinline__ uint32_t crc_mul(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
// x^(8 * len) mod P from the table x2n[k] = x^(8 * 2^k)
__device__ __forceinline__ uint32_t crc_xpow(const uint32_t *x2n, uint32_t len)
{
    uint32_t p = 1u << 31;                                     // x^0
    for (int k = 0; len; ++k, len >>= 1) if (len & 1) p = crc_mul(x2n[k], p);
    return p;
}

__global__ void __launch_bounds__(kGzThreads)
gz_compress_kernel(const uint8_t *__restrict__ raw, unsigned long long n_raw, const GzDeviceTables T,
                   const uint32_t *__restrict__ crc_tab /* [4][256] then x2n[32] */,
                   uint8_t *__restrict__ slots, unsigned long long *__restrict__ sizes)
{
    __shared__ uint32_t s_code[257];
    __shared__ uint32_t s_crc[4 * 256];
    __shared__ uint32_t s_x2n[32];
    __shared__ uint32_t s_scan[kGzThreads];                    // bit counts -> exclusive offsets
    __shared__ uint32_t s_tail[kGzThreads];                    // bits a chunk leaves in the word it shares with the next chunk
    __shared__ uint32_t s_v[kGzThreads], s_l[kGzThreads];      // CRC combine: value, length
    const int t = threadIdx.x;
    for (int j = t; j < 257; j += kGzThreads) s_code[j] = T.code[j];
    for (int j = t; j < 1024; j += kGzThreads) s_crc[j] = crc_tab[j];
    if (t < 32) s_x2n[t] = crc_tab[1024 + t];
    __syncthreads();

    const unsigned long long m = blockIdx.x, base = m * kGzMemberRaw;
    const int mlen = (int)min((unsigned long long)kGzMemberRaw, n_raw - base);
    const int c0 = t * kGzChunk, clen = max(0, min(kGzChunk, mlen - c0));
    const uint8_t *src = raw + base + c0;                      // 256-byte aligned: the stream buffers are, and so is c0
    uint32_t *out = reinterpret_cast<uint32_t *>(slots + m * (unsigned long long)kGzSlotStride);

    // ---- pass 1: bits of the chunk and its CRC-32 ---------------------------------------------------------------------
    uint32_t bits = 0, crc = 0xFFFFFFFFu;
    for (int x = 0; x < clen; x += 16) {
        const uint4 v = *reinterpret_cast<const uint4 *>(src + x);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (x + 4 * q + 4 <= clen) {
                bits += (s_code[w[q] & 0xFF] & 15u) + (s_code[(w[q] >> 8) & 0xFF] & 15u) + (s_code[(w[q] >> 16) & 0xFF] & 15u) + (s_code[w[q] >> 24] & 15u);
                const uint32_t y = crc ^ w[q];
                crc = s_crc[3 * 256 + (y & 0xFF)] ^ s_crc[2 * 256 + ((y >> 8) & 0xFF)] ^ s_crc[256 + ((y >> 16) & 0xFF)] ^ s_crc[y >> 24];
            } else {
                for (int b = 0; b < 4 && x + 4 * q + b < clen; ++b) {
                    const uint32_t ch = (w[q] >> (8 * b)) & 0xFFu;
                    bits += s_code[ch] & 15u;
                    crc = s_crc[(crc ^ ch) & 0xFF] ^ (crc >> 8);
                }
            }
        }
    }
    crc ^= 0xFFFFFFFFu;
    const int last = (mlen - 1) / kGzChunk;                    // last non-empty chunk (mlen >= 1)
    if (t == last) bits += s_code[256] & 15u;                  // end-of-block code
    // exclusive scan of the bit counts
    s_scan[t] = bits;
    __syncthreads();
    for (int o = 1; o < kGzThreads; o <<= 1) {
        const uint32_t add = t >= o ? s_scan[t - o] : 0u;
        __syncthreads();
        s_scan[t] += add;
        __syncthreads();
    }
    const uint32_t total_bits = T.prefix_bits + s_scan[kGzThreads - 1];
    const uint32_t start = T.prefix_bits + s_scan[t] - bits;   // first bit of this chunk in the member

    // ---- CRC-32 of the member: combine the chunk CRCs, crc(A || B) = crc(A) * x^(8|B|) + crc(B) --------------------------------
    s_v[t] = clen ? crc : 0u; s_l[t] = (uint32_t)clen;
    __syncthreads();
    for (int lvl = 0; lvl < 8; ++lvl) {
        const int span = 1 << lvl;
        if ((t & (2 * span - 1)) == 0) {
            const uint32_t lb = s_l[t + span];
            if (lb) {
                // full right halves have 256 * 2^lvl bytes: x^(8 * 2^(8 + lvl)) is a table entry
                const uint32_t xp = lb == (uint32_t)(kGzChunk << lvl) ? s_x2n[8 + lvl] : crc_xpow(s_x2n, lb);
                s_v[t] = crc_mul(xp, s_v[t]) ^ s_v[t + span];
                s_l[t] += lb;
            }
        }
        __syncthreads();
    }
    const uint32_t member_crc = s_v[0];

    // ---- pass 2: pack the codes, LSB first ----------------------------------------------------------------------------------
    // prefix: whole words first; its partial last word is OR'ed into chunk 0's first word below
    const uint32_t pw = T.prefix_bits >> 5;
    for (uint32_t j = t; j < pw; j += kGzThreads) out[j] = reinterpret_cast<const uint32_t *>(T.prefix)[j];
    uint32_t w_idx = start >> 5, head_idx = w_idx, head = 0, tail = 0;
    bool have_head = false;
    if (t <= last) {
        uint64_t acc = 0;
        int nb = (int)(start & 31u);
        auto push = [&](uint32_t code) {
            acc |= (uint64_t)(code >> 4) << nb;
            nb += (int)(code & 15u);
            if (nb >= 32) {
                if (!have_head) { head = (uint32_t)acc; have_head = true; } else out[w_idx] = (uint32_t)acc;
                ++w_idx; acc >>= 32; nb -= 32;
            }
        };
        for (int x = 0; x < clen; x += 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + x);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int b = 0; b < 16; ++b) if (x + b < clen) push(s_code[(w[b >> 2] >> (8 * (b & 3))) & 0xFFu]);
        }
        if (t == last) push(s_code[256]);
        tail = (uint32_t)acc;                                   // nb < 32 bits left in word w_idx
    }
    s_tail[t] = tail;
    __syncthreads();
    if (t <= last) {
        // bits that precede this chunk inside its first word: the prefix's partial word, or the previous chunk's tail
        uint32_t before = 0;
        if (start & 31u) before = t == 0 ? reinterpret_cast<const uint32_t *>(T.prefix)[pw] : s_tail[t - 1];
        if (have_head) out[head_idx] = head | before;
        else tail |= before;                                    // only possible for a short last chunk
        if (t == last) {
            // final partial word, then CRC-32 and ISIZE at the next byte boundary
            uint8_t *bytes = reinterpret_cast<uint8_t *>(out);
            const uint32_t end_byte = (total_bits + 7) >> 3;
            for (uint32_t x = w_idx * 4; x < end_byte; ++x) bytes[x] = (uint8_t)(tail >> (8 * (x - w_idx * 4)));
            for (int b = 0; b < 4; ++b) bytes[end_byte + b] = (uint8_t)(member_crc >> (8 * b));
            for (int b = 0; b < 4; ++b) bytes[end_byte + 4 + b] = (uint8_t)((uint32_t)mlen >> (8 * b));
            sizes[m] = end_byte + 8;
        }
    }
}

// members -> contiguous stream; offs = exclusive scan of sizes
__global__ void __launch_bounds__(256)
gz_compact_kernel(const uint8_t *__restrict__ slots, const unsigned long long *__restrict__ sizes,
                  const unsigned long long *__restrict__ offs, uint8_t *__restrict__ out, unsigned long long cap)
{
    const unsigned long long m = blockIdx.x;
    const uint8_t *src = slots + m * (unsigned long long)kGzSlotStride;
    uint8_t *dst = out + offs[m];
    const uint32_t n = (uint32_t)sizes[m];
    // a code fitted to the first batch can expand later data up to 15 bits per byte: members that would not fit the stream's
    // buffer are not written (the host sees the total, reports the overflow and nothing leaves the device)
    if (offs[m] + n > cap) return;
    // aligned middle of the destination with 4-byte stores, source read byte-wise shifted (slots are 4-byte aligned)
    const uint32_t lead = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    for (uint32_t x = threadIdx.x; x < min(lead, n); x += blockDim.x) dst[x] = src[x];
    if (n > lead) {
        const uint32_t nw = (n - lead) >> 2;
        const uint32_t sh = (lead & 3) * 8;
        const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + lead);
        for (uint32_t j = threadIdx.x; j < nw; j += blockDim.x) {
            // destination word j holds source bytes [lead + 4j, lead + 4j + 4)
            const uint32_t a = s32[(lead >> 2) + j], b = s32[(lead >> 2) + j + 1];
            d32[j] = sh ? __funnelshift_r(a, b, sh) : a;
        }
        for (uint32_t x = lead + nw * 4 + threadIdx.x; x < n; x += blockDim.x) dst[x] = src[x];
    }
}

}  // namespace dwg
