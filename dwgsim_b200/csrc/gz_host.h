// gz_host.h -- host side of the device gzip writer: length-limited Huffman codes, the constant member prefix
// (gzip header + dynamic-Huffman block header), CRC-32 constants, and a CPU encoder of exactly the same member
// format (used by the CPU tests to validate the tables with zlib's inflate).
//
// Member format (RFC 1952 / RFC 1951): 10-byte gzip header, ONE deflate block (BFINAL=1, BTYPE=2) that codes every
// byte as a literal with a per-stream Huffman code, the end-of-block code, zero padding to a byte boundary, CRC-32
// and ISIZE.  FASTQ gains little from LZ77 matches (random bases and qualities), so literal-only coding with a code
// fitted to the stream's symbol statistics gets within a few percent of `gzip -6` at a fraction of the work.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace dwg {

constexpr int kGzMemberRaw = 1 << 16;          // raw bytes per gzip member
constexpr int kGzChunk = 256;                  // raw bytes per thread
constexpr int kGzSlotStride = 2 * kGzMemberRaw + 1024;   // worst case: 15 bits per byte + prefix + trailer

struct GzTables {
    uint32_t code[257];                        // (bit-reversed code << 4) | length, symbols 0..255 and 256 = end of block
    std::vector<uint8_t> prefix;               // gzip header + block header, bit-packed LSB first
    uint32_t prefix_bits = 0;                  // number of valid bits in prefix
};

// Huffman code lengths for freq[0..n) limited to maxbits (every symbol with freq > 0 gets a code)
inline void huffman_lengths(const uint64_t *freq, int n, int maxbits, uint8_t *len)
{
    std::vector<uint64_t> f(freq, freq + n);
    for (;;) {
        struct Node { uint64_t w; int left, right; };
        std::vector<Node> nodes;
        std::vector<int> live;
        for (int i = 0; i < n; ++i) { len[i] = 0; if (f[i]) { nodes.push_back({f[i], -1 - i, -1}); live.push_back((int)nodes.size() - 1); } }
        if (live.empty()) return;
        if (live.size() == 1) { len[-1 - nodes[live[0]].left] = 1; return; }
        while (live.size() > 1) {                      // O(n^2) merge of the two lightest nodes; n <= 286
            auto lightest = [&]() {
                size_t b = 0;
                for (size_t i = 1; i < live.size(); ++i) if (nodes[live[i]].w < nodes[live[b]].w) b = i;
                int id = live[b];
                live.erase(live.begin() + (long)b);
                return id;
            };
            const int a = lightest(), b = lightest();
            nodes.push_back({nodes[a].w + nodes[b].w, a, b});
            live.push_back((int)nodes.size() - 1);
        }
        int maxlen = 0;
        std::vector<std::pair<int, int>> stack{{live[0], 0}};
        while (!stack.empty()) {
            auto [id, d] = stack.back();
            stack.pop_back();
            if (nodes[id].right < 0 && nodes[id].left < 0) { len[-1 - nodes[id].left] = (uint8_t)d; maxlen = std::max(maxlen, d); }
            else { stack.push_back({nodes[id].left, d + 1}); stack.push_back({nodes[id].right, d + 1}); }
        }
        if (maxlen <= maxbits) return;
        for (auto &x : f) if (x) x = (x + 1) >> 1;      // flatten the distribution and rebuild
    }
}

// canonical codes (RFC 1951 3.2.2), returned bit-reversed so they can be OR'ed into an LSB-first bit stream
inline void canonical_codes(const uint8_t *len, int n, uint32_t *code_rev)
{
    uint32_t bl_count[16] = {0}, next[16] = {0};
    for (int i = 0; i < n; ++i) bl_count[len[i]]++;
    bl_count[0] = 0;
    uint32_t c = 0;
    for (int b = 1; b < 16; ++b) { c = (c + bl_count[b - 1]) << 1; next[b] = c; }
    for (int i = 0; i < n; ++i) {
        uint32_t v = 0;
        if (len[i]) { uint32_t x = next[len[i]]++; for (int b = 0; b < len[i]; ++b) v |= ((x >> b) & 1u) << (len[i] - 1 - b); }
        code_rev[i] = v;
    }
}

struct BitWriter {
    std::vector<uint8_t> bytes;
    uint64_t nbits = 0;
    void put(uint32_t v, int n)
    {
        for (int i = 0; i < n; ++i) {
            if ((nbits & 7) == 0) bytes.push_back(0);
            bytes.back() |= (uint8_t)(((v >> i) & 1u) << (nbits & 7));
            ++nbits;
        }
    }
};

// build the per-stream tables from a byte histogram (every literal gets a code, so any byte stays encodable)
inline GzTables gz_build_tables(const uint64_t hist[256])
{
    GzTables t;
    uint64_t freq[257];
    for (int i = 0; i < 256; ++i) freq[i] = std::max<uint64_t>(hist[i], 1);
    uint64_t total = 0;
    for (int i = 0; i < 256; ++i) total += freq[i];
    freq[256] = std::max<uint64_t>(total / kGzMemberRaw, 1);       // one end-of-block per member
    uint8_t len[257];
    huffman_lengths(freq, 257, 15, len);
    uint32_t code[257];
    canonical_codes(len, 257, code);
    for (int i = 0; i < 257; ++i) t.code[i] = (code[i] << 4) | len[i];
    // distance tree: two codes of one bit each (what zlib emits for a block without matches)
    const uint8_t dlen[2] = {1, 1};
    // code-length alphabet: every literal/length and distance length is sent verbatim (no run-length symbols)
    uint64_t clfreq[19] = {0};
    for (int i = 0; i < 257; ++i) clfreq[len[i]]++;
    clfreq[1] += 2;
    uint8_t cllen[19];
    huffman_lengths(clfreq, 19, 7, cllen);
    uint32_t clcode[19];
    canonical_codes(cllen, 19, clcode);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int ncl = 19;
    while (ncl > 4 && cllen[order[ncl - 1]] == 0) --ncl;
    BitWriter w;
    const uint8_t gz[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 3};
    for (uint8_t b : gz) w.put(b, 8);
    w.put(1, 1);                 // BFINAL
    w.put(2, 2);                 // BTYPE = dynamic Huffman
    w.put(257 - 257, 5);         // HLIT
    w.put(2 - 1, 5);             // HDIST
    w.put((uint32_t)(ncl - 4), 4);
    for (int i = 0; i < ncl; ++i) w.put(cllen[order[i]], 3);
    for (int i = 0; i < 257; ++i) w.put(clcode[len[i]], cllen[len[i]]);
    for (int i = 0; i < 2; ++i) w.put(clcode[dlen[i]], cllen[dlen[i]]);
    t.prefix = w.bytes;
    t.prefix_bits = (uint32_t)w.nbits;
    return t;
}

// ---- CRC-32 (IEEE 802.3, reflected), tables for slicing-by-4 and the GF(2) constants used to combine chunk CRCs ------
struct Crc32Tables {
    uint32_t t[4][256];
    uint32_t x2n[32];            // x^(8 * 2^k) mod P, k = 0..31, in zlib's reflected representation
    Crc32Tables()
    {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; t[0][i] = c; }
        for (uint32_t i = 0; i < 256; ++i) for (int s = 1; s < 4; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
        uint32_t p = 1u << 30;                         // x^1
        p = mul(p, p); p = mul(p, p); p = mul(p, p);   // x^8
        for (int k = 0; k < 32; ++k) { x2n[k] = p; p = mul(p, p); }
    }
    static uint32_t mul(uint32_t a, uint32_t b)       // a(x) * b(x) mod P, reflected (zlib's multmodp)
    {
        uint32_t m = 1u << 31, p = 0;
        for (;;) {
            if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
            m >>= 1;
            b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
        }
        return p;
    }
    uint32_t shift(uint32_t crc, uint64_t len) const  // crc of A  ->  contribution of A in crc(A || B), |B| = len
    {
        for (int k = 0; len; ++k, len >>= 1) if (len & 1) crc = mul(x2n[k], crc);
        return crc;
    }
    uint32_t crc(const uint8_t *p, size_t n) const
    {
        uint32_t c = 0xFFFFFFFFu;
        for (size_t i = 0; i < n; ++i) c = t[0][(c ^ p[i]) & 0xFF] ^ (c >> 8);
        return c ^ 0xFFFFFFFFu;
    }
};

// CPU encoder of the member format (tests): appends the gzip members of data[0..n) to out
inline void gz_encode_host(const GzTables &t, const Crc32Tables &ct, const uint8_t *data, size_t n, std::vector<uint8_t> &out)
{
    for (size_t off = 0; off < n || (n == 0 && off == 0); off += kGzMemberRaw) {
        const size_t m = std::min<size_t>(kGzMemberRaw, n - off);
        BitWriter w;
        w.bytes = t.prefix;
        w.nbits = t.prefix_bits;
        for (size_t i = 0; i < m; ++i) w.put(t.code[data[off + i]] >> 4, (int)(t.code[data[off + i]] & 15));
        w.put(t.code[256] >> 4, (int)(t.code[256] & 15));
        out.insert(out.end(), w.bytes.begin(), w.bytes.end());
        const uint32_t c = ct.crc(data + off, m), isz = (uint32_t)m;
        for (int b = 0; b < 4; ++b) out.push_back((uint8_t)(c >> (8 * b)));
        for (int b = 0; b < 4; ++b) out.push_back((uint8_t)(isz >> (8 * b)));
        if (n == 0) break;
    }
}

}  // namespace dwg
