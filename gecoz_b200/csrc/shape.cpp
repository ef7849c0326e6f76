// Host-side, byte-defining part of the path: code lengths, canonical codes, tree order and the
// serialized length table of one block (what the reference computes in HSWTShape).
//
// Mirrors (semantics, tie-breaks and output bytes; the structure is our own):
//   algo/huffman/HuffmanEncodeTable.java:48-111     which pair is merged next -> code LENGTHS
//   algo/deflate/DeflateEncodeTable.java:63-173     length limiting, canonical bit-reversed codes
//   algo/deflate/DeflateLengthsTable.java:82-207    RFC 1951 3.2.7-style table (HCLEN, RLE 16/17/18)
//   algo/tree/HSWTShape.java:55-115                 node sizes, serialized size, write/read
//   algo/tree/HuffmanShapedWaveletTree.java:165-182 node order in the file (pre-order, 0-branch first)
#include "gcz_host.h"

#include <algorithm>
#include <array>
#include <cstring>
#include <limits>
#include <vector>

namespace gcz {

// ---- code lengths ------------------------------------------------------------------------------
// The reference repeatedly scans the weights for the two smallest positive ones.  The scan order
// defines the tie-breaks: a strictly smaller weight takes slot 1 and pushes the previous holder to
// slot 2; an equal weight never displaces slot 1.  The pair is merged INTO slot 2's index.
static std::vector<int> huffman_lengths(const std::vector<int64_t>& weights_in) {
    const int n = (int)weights_in.size();
    std::vector<int64_t> w(weights_in);
    std::vector<int> len(n, 0);
    // leaves currently hanging under the subtree rooted at index i (singly linked, tail pointer kept)
    std::vector<int> next_leaf(n, -1), tail(n);
    for (int i = 0; i < n; i++) tail[i] = i;

    for (int round = 1; round < n; round++) {
        int a = 0, b = 0;
        int64_t wa = std::numeric_limits<int64_t>::max(), wb = wa;
        for (int j = 0; j < n; j++) {
            const int64_t f = w[j];
            if (f <= 0) continue;
            if (f < wa)      { b = a; wb = wa; a = j; wa = f; }
            else if (f < wb) { b = j; wb = f; }
        }
        if (wb == std::numeric_limits<int64_t>::max()) {
            if (round == 1) len[a] = 1;          // a single distinct symbol still gets one bit
            break;
        }
        w[a] = std::numeric_limits<int64_t>::min();
        w[b] = wa + wb;
        for (int i = a; i >= 0; i = next_leaf[i]) len[i]++;
        for (int i = b; i >= 0; i = next_leaf[i]) len[i]++;
        next_leaf[tail[b]] = a;                  // b's subtree absorbs a's leaves
        tail[b] = tail[a];
    }
    return len;
}

// Length limiting exactly as DeflateEncodeTable.restrict_lengths (a heuristic, not package-merge):
// leaves are kept in a list sorted by (length, weight, index); leaves are pushed down a level from the
// top of that order until enough slots are freed, then pulled back up from the bottom if over-shot.
static void limit_lengths(const std::vector<int64_t>& weights, std::vector<int>& len, int max_bits) {
    const int n = (int)len.size();
    int64_t total = 0;
    std::vector<int> per_len(std::max(n, 64) + 1, 0);
    for (int i = 0; i < n; i++) if (len[i] > 0) { total += len[i]; per_len[len[i]]++; }
    if (total <= 1) return;

    int32_t free_nodes = 1;
    for (int l = 1; l <= max_bits && free_nodes > 0; l++) {
        free_nodes = (int32_t)((uint32_t)free_nodes << 1);
        free_nodes -= per_len[l];
    }
    if (free_nodes <= 0) return;                  // nothing is deeper than max_bits

    free_nodes = -free_nodes;
    for (int i = 0; i < n; i++) if (len[i] > max_bits) { len[i] = max_bits; free_nodes++; }

    struct Leaf { int64_t key; };                 // (length << 48) | (weight << 16) | index, as in the reference
    std::vector<int64_t> order(n);
    for (int i = 0; i < n; i++)
        order[i] = (int64_t)(((uint64_t)(int64_t)len[i] << 48) | ((uint64_t)weights[i] << 16) | (uint64_t)i);
    std::sort(order.begin(), order.end());
    auto level_of = [](int64_t v) { return (int)(int8_t)((uint64_t)v >> 48); };
    auto with_level = [](int64_t v, int l) { return (int64_t)(((uint64_t)v & 0xFF00FFFFFFFFFFFFULL) | ((uint64_t)l << 48)); };

    do {
        bool enough = false;
        for (int start = max_bits - 1; start > 0 && !enough; start--) {
            for (int level = start; level < max_bits && !enough; level++) {
                for (int j = 0; j < n; j++) {
                    if (level_of(order[j]) == level) {
                        order[j] = with_level(order[j], level + 1);
                        free_nodes -= 1 << (max_bits - 1 - level);
                        if (free_nodes <= 0) { enough = true; break; }
                    }
                }
            }
        }
        for (int level = max_bits; free_nodes < 0 && level > 0; level--) {
            for (int i = n - 1; free_nodes < 0 && i >= 0; i--) {
                if (level_of(order[i]) == level) {
                    order[i] = with_level(order[i], level - 1);
                    free_nodes += 1 << (max_bits - level);
                }
            }
        }
    } while (free_nodes != 0);

    for (int i = 0; i < n; i++) len[(int)(order[i] & 0xFFFF)] = level_of(order[i]);
}

static uint32_t bit_reverse16(uint32_t v) {
    v = ((v & 0x5555u) << 1) | ((v >> 1) & 0x5555u);
    v = ((v & 0x3333u) << 2) | ((v >> 2) & 0x3333u);
    v = ((v & 0x0F0Fu) << 4) | ((v >> 4) & 0x0F0Fu);
    return ((v >> 8) | (v << 8)) & 0xFFFFu;
}

// RFC 1951 canonical codes, stored root-first (bit j = branch at depth j).
static bool canonical_codes(const std::vector<int>& len, int max_bits, std::vector<int>& code) {
    const int n = (int)len.size();
    std::vector<int> per_len(max_bits + 2, 0), next(max_bits + 2, 0);
    for (int i = 0; i < n; i++) {
        if (len[i] > max_bits) return false;
        if (len[i] > 0) per_len[len[i]]++;
    }
    for (int bits = 1, c = 0; bits <= max_bits; bits++) {
        c = (c + per_len[bits - 1]) << 1;
        next[bits] = c;
    }
    code.assign(n, 0);
    for (int i = 0; i < n; i++) {
        if (len[i] == 0) continue;
        code[i] = (int)(int16_t)(bit_reverse16((uint32_t)next[len[i]]) >> (16 - len[i]));
        next[len[i]]++;
    }
    return true;
}

bool deflate_code(const std::vector<int64_t>& weights, int max_bits, std::vector<int>& len, std::vector<int>& code) {
    len = huffman_lengths(weights);
    limit_lengths(weights, len, max_bits);
    return canonical_codes(len, max_bits, code);
}

// ---- LSB-first bit sink / source (io/AbstractBitStream.java packs the same way: C.1 KAT) ------------
struct BitSink {
    std::vector<uint8_t> bytes;
    uint64_t acc = 0; int used = 0;
    void put(uint32_t v, int nbits) {
        for (int i = 0; i < nbits; i++) {
            acc |= (uint64_t)((v >> i) & 1u) << used;
            if (++used == 8) { bytes.push_back((uint8_t)acc); acc = 0; used = 0; }
        }
    }
    void finish() { if (used) { bytes.push_back((uint8_t)acc); acc = 0; used = 0; } }
    int64_t bit_count() const { return (int64_t)bytes.size() * 8 + used; }
};

struct BitSource {
    const uint8_t* p; int64_t nbytes; int64_t bitpos = 0;
    bool eof = false;
    uint32_t get(int nbits) {
        uint32_t v = 0;
        for (int i = 0; i < nbits; i++) {
            const int64_t byte = bitpos >> 3;
            if (byte >= nbytes) { eof = true; return v; }
            v |= (uint32_t)((p[byte] >> (bitpos & 7)) & 1u) << i;
            bitpos++;
        }
        return v;
    }
};

// ---- the serialized length table ---------------------------------------------------------------
static const int kClOrder[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };

// One walk over the 256 lengths produces the RLE token stream; `emit(sym, extra_value, extra_bits)`.
// Run handling follows DeflateLengthsTable.write :94-124: a run is flushed when the length changes or at
// the last index, which is always emitted as a literal; a leading run of zeros has no literal in front.
template <class Emit>
static void rle_tokens(const int8_t* lens, int n, Emit emit) {
    int cur = 0, run = 0;
    for (int i = 0; i < n; i++) {
        if (cur != lens[i] || i == n - 1) {
            while (run >= 3) {
                if (cur != 0)       { run -= 3; emit(16, std::min(run, 3), 2);   run -= 3; }
                else if (run <= 10) { run -= 3; emit(17, std::min(run, 7), 3);   run -= 7; }
                else                { run -= 11; emit(18, std::min(run, 127), 7); run -= 127; }
            }
            while (run-- > 0) emit(cur, 0, 0);
            cur = lens[i];
            emit(cur, 0, 0);
            run = 0;
        } else {
            run++;
        }
    }
}

struct LengthTableCode {
    std::vector<int> len, code;
    int last_used = -1;          // index into kClOrder of the last code-length symbol in use ("hclen")
};

static bool length_table_code(const int8_t* lens, int max_bits, LengthTableCode& out) {
    std::vector<int64_t> freq(19, 0);
    rle_tokens(lens, 256, [&](int sym, int, int) { freq[sym]++; });
    out.last_used = -1;
    for (int i = 18; i >= 0; i--) if (freq[kClOrder[i]] > 0) { out.last_used = i; break; }
    return deflate_code(freq, max_bits, out.len, out.code);
}

// DeflateLengthsTable.length(): sized with an UNRESTRICTED (15-bit) code-length code.
static int64_t length_table_bits(const int8_t* lens) {
    LengthTableCode c;
    if (!length_table_code(lens, 15, c)) return -1;
    int64_t bits = 7 + 3 * (int64_t)c.last_used;
    rle_tokens(lens, 256, [&](int sym, int, int extra_bits) { bits += c.len[sym] + extra_bits; });
    return bits;
}

// DeflateLengthsTable.write(): emitted with a 7-bit-limited code-length code.
static bool length_table_write(const int8_t* lens, BitSink& out) {
    LengthTableCode c;
    if (!length_table_code(lens, 7, c)) return false;
    out.put((uint32_t)(c.last_used - 3) & 15u, 4);
    for (int i = 0; i <= c.last_used; i++) out.put((uint32_t)c.len[kClOrder[i]], 3);
    rle_tokens(lens, 256, [&](int sym, int extra, int extra_bits) {
        out.put((uint32_t)c.code[sym], c.len[sym]);
        if (extra_bits) out.put((uint32_t)extra, extra_bits);
    });
    return true;
}

// ---- tree ------------------------------------------------------------------------------------------
// Pre-order walk over the prefix tree of the codes; an internal node is named after the left-most leaf
// of its 1-subtree, which is what decode.getSymbol(prefix | 1 << depth) returns in the reference.
struct Trie {
    struct Node { int child[2] = { -1, -1 }; int symbol = -1; };
    std::vector<Node> nodes;
    Trie() { nodes.emplace_back(); }
    bool insert(int symbol, int code, int len) {
        int v = 0;
        for (int j = 0; j < len; j++) {
            if (nodes[v].symbol >= 0) return false;               // not prefix free
            const int b = (code >> j) & 1;
            if (nodes[v].child[b] < 0) { nodes[v].child[b] = (int)nodes.size(); nodes.emplace_back(); }
            v = nodes[v].child[b];
        }
        if (nodes[v].child[0] >= 0 || nodes[v].child[1] >= 0 || nodes[v].symbol >= 0) return false;
        nodes[v].symbol = symbol;
        return true;
    }
    int leftmost_leaf(int v) const {
        while (v >= 0 && nodes[v].symbol < 0) v = nodes[v].child[0] >= 0 ? nodes[v].child[0] : nodes[v].child[1];
        return v < 0 ? -1 : nodes[v].symbol;
    }
};

static bool build_tree(gcz_shape* s) {
    Trie trie;
    int present = 0, only = -1;
    for (int i = 0; i < 256; i++) {
        if (s->bit_lengths[i] > 0) {
            present++; only = i;
            if (!trie.insert(i, (uint16_t)s->codes[i], s->bit_lengths[i])) return false;
        }
    }
    s->n_nodes = 0;
    if (present == 0) return true;
    if (present == 1) {
        // Single symbol: code 0 of length 1 after canonical remapping; the reference names the root through
        // lookup entry 1, which is empty and therefore decodes to symbol 0.  Only symbol 0 itself makes a
        // usable tree there (any other symbol hits a null node in fill()): mirror that restriction.
        if (only != 0) return false;
        s->node_name[0] = 0; s->node_depth[0] = 0; s->node_prefix[0] = 0;
        s->n_nodes = 1;
        return true;
    }
    struct Item { int v, depth, prefix; };
    std::vector<Item> stack;
    stack.push_back({ 0, 0, 0 });
    while (!stack.empty()) {
        const Item it = stack.back(); stack.pop_back();
        const Trie::Node& nd = trie.nodes[it.v];
        if (nd.symbol >= 0) continue;
        if (nd.child[0] < 0 || nd.child[1] < 0) return false;     // a Huffman tree is full
        const int k = s->n_nodes++;
        s->node_name[k] = trie.leftmost_leaf(nd.child[1]);
        s->node_depth[k] = it.depth;
        s->node_prefix[k] = it.prefix;
        stack.push_back({ nd.child[1], it.depth + 1, it.prefix | (1 << it.depth) });   // visited second
        stack.push_back({ nd.child[0], it.depth + 1, it.prefix });                      // visited first
    }
    return true;
}

int64_t ranked_bytes(int64_t len) {
    return (int64_t)(((uint64_t)(len - 1) >> 16) * 6 + ((uint64_t)(len - 1) >> 9) * 2 + ((uint64_t)(len + 7) >> 3));
}

int64_t index_size(int64_t n, int sampling_factor) {
    const int64_t m = (n + ((int64_t)1 << sampling_factor) - 1) >> sampling_factor;
    const int levels = m == 0 ? 0 : 64 - __builtin_clzll((uint64_t)m);
    return ranked_bytes(m) * levels + ranked_bytes(n);
}

int shape_from_counts(const int64_t counts[256], gcz_shape* s) {
    std::memset(s, 0, sizeof(*s));
    std::vector<int64_t> w(counts, counts + 256);
    for (int i = 0; i < 256; i++) if (w[i] < 0) return fail(GCZ_E_ARG, "negative symbol count");
    std::vector<int> len, code;
    if (!deflate_code(w, 15, len, code)) return fail(GCZ_E_RANGE, "code length above 15 bits");
    int64_t total = 0;
    for (int i = 0; i < 256; i++) {
        s->bit_lengths[i] = (int8_t)len[i];
        s->codes[i] = (int16_t)code[i];
        total += counts[i];
    }
    if (total > std::numeric_limits<int32_t>::max()) return fail(GCZ_E_RANGE, "block above 2^31-1 symbols");
    if (!build_tree(s)) return fail(GCZ_E_RANGE, "degenerate alphabet: the reference cannot build this tree");

    // node sizes: every symbol contributes its count to each node on its path
    for (int k = 0; k < s->n_nodes; k++) {
        const int d = s->node_depth[k], prefix = s->node_prefix[k];
        int64_t bits = 0;
        for (int i = 0; i < 256; i++) {
            if (counts[i] > 0 && s->bit_lengths[i] > d && ((uint16_t)s->codes[i] & ((1 << d) - 1)) == prefix) bits += counts[i];
        }
        s->node_bits[k] = bits;
    }
    const int64_t tbits = length_table_bits(s->bit_lengths);
    if (tbits < 0) return fail(GCZ_E_INTERNAL, "length table sizing failed");
    s->table_bytes = (tbits + 7) >> 3;
    // the reference sizes the table with a 15-bit code-length code but writes it with a 7-bit one; when the
    // two disagree its own file is inconsistent, so refuse instead of guessing
    BitSink probe;
    if (!length_table_write(s->bit_lengths, probe)) return fail(GCZ_E_RANGE, "code-length code above 7 bits");
    probe.finish();
    if ((int64_t)probe.bytes.size() != s->table_bytes) return fail(GCZ_E_RANGE, "length table size mismatch (reference quirk)");

    int64_t off = s->table_bytes;
    for (int k = 0; k < s->n_nodes; k++) {
        s->node_offset[k] = off;
        if (s->node_bits[k] > 0) off += ranked_bytes(s->node_bits[k]);
    }
    s->length = total;
    s->size = off;
    return GCZ_OK;
}

int64_t shape_write(const gcz_shape* s, uint8_t* out, int64_t cap) {
    BitSink sink;
    if (!length_table_write(s->bit_lengths, sink)) return fail(GCZ_E_RANGE, "code-length code above 7 bits");
    sink.finish();
    if ((int64_t)sink.bytes.size() > cap) return fail(GCZ_E_ARG, "shape buffer too small");
    std::memcpy(out, sink.bytes.data(), sink.bytes.size());
    return (int64_t)sink.bytes.size();
}

// Canonical decoder for the code-length code: walk bit by bit (codes are stored root-first).
static int decode_symbol(BitSource& in, const std::vector<int>& len, const std::vector<int>& code) {
    int acc = 0;
    for (int depth = 1; depth <= 15; depth++) {
        acc |= (int)in.get(1) << (depth - 1);
        if (in.eof) return -1;
        for (size_t s = 0; s < len.size(); s++) if (len[s] == depth && code[s] == acc) return (int)s;
    }
    return -1;
}

int shape_read(const uint8_t* body, int64_t body_len, gcz_shape* s) {
    std::memset(s, 0, sizeof(*s));
    BitSource in{ body, body_len };
    const int hclen = (int)in.get(4) + 4;
    std::vector<int> cl_len(19, 0), cl_code;
    for (int i = 0; i < hclen; i++) cl_len[kClOrder[i]] = (int)in.get(3);
    if (in.eof || !canonical_codes(cl_len, 15, cl_code)) return fail(GCZ_E_FORMAT, "bad shape table header");
    int last = 0;
    for (int i = 0; i < 256;) {
        const int sym = decode_symbol(in, cl_len, cl_code);
        if (sym < 0) return fail(GCZ_E_FORMAT, "bad shape table code");
        if (sym <= 15) { s->bit_lengths[i++] = (int8_t)(last = sym); }
        else if (sym == 16) {
            const int rep = (int)in.get(2) + 3;
            if (i + rep > 256) return fail(GCZ_E_FORMAT, "shape table overruns 256 symbols");
            for (int j = 0; j < rep; j++) s->bit_lengths[i++] = (int8_t)last;
        }
        else if (sym == 17) i += (int)in.get(3) + 3;
        else                i += (int)in.get(7) + 11;
        if (in.eof) return fail(GCZ_E_FORMAT, "truncated shape table");
    }
    s->table_bytes = (in.bitpos + 7) >> 3;
    std::vector<int> len(256), code;
    for (int i = 0; i < 256; i++) len[i] = s->bit_lengths[i];
    if (!canonical_codes(len, 15, code)) return fail(GCZ_E_FORMAT, "shape table holds a length above 15");
    for (int i = 0; i < 256; i++) s->codes[i] = (int16_t)code[i];
    if (!build_tree(s)) return fail(GCZ_E_FORMAT, "shape table is not a full prefix code");
    return GCZ_OK;
}

}  // namespace gcz
