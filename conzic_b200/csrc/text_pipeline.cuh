// BERT ids -> CLIP ids without leaving the device: what the reference does per Gibbs step with two host tokenizers,
//   texts = tokenizer.batch_decode(ids, skip_special_tokens=True)          (gen_utils.py:75)
//   clip_ids = CLIPTokenizer(texts, padding=True, truncation=True, ...)    (clip/clip.py:71-72)
// restated as one function over per-token byte tables, so that vocabularies with '##' word pieces (every real BERT
// vocabulary) need no string round trip:
//   * WordPiece decoder of the `tokenizers` library (decoders/wordpiece.rs, prefix "##", cleanup = true): special ids
//     dropped; the first kept token verbatim; a '##' piece glued to what precedes it; every other token preceded by a
//     space unless the per-token clean-up removes it (token starts with . ? ! , or n't 'm 's 've 're);
//   * CLIP normaliser (NFC, whitespace collapse, lowercase: applied per token on the host when the byte table is
//     built) and pre-tokeniser regex  's|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+  as a scanner over
//     (byte, character class) pairs; the ByteLevel stage's own regex never splits these pieces further;
//   * byte-level BPE with end-of-word suffix "</w>" (models/bpe: lowest merge rank first, leftmost first), merges
//     looked up in an open-addressing hash table keyed by the pair of symbol ids; a pre-token that is exactly one
//     whole "simple" BERT token takes its precomputed ids instead (the common case: a whole word between spaces).
// The functions are __host__ __device__: the CUDA kernels in text_ops.cu and the CPU test harness
// (tests/text_host_harness.cu, checked against the real transformers tokenizer classes) compile this same code.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CZ_HD __host__ __device__ __forceinline__
#else
#define CZ_HD inline
#endif

namespace conzic {

enum { TXT_CLS_OTHER = 0, TXT_CLS_LETTER = 1, TXT_CLS_NUMBER = 2, TXT_CLS_SPACE = 3, TXT_CHAR_START = 4 };
enum { TXT_FLAG_PIECE = 1, TXT_FLAG_GLUE = 2, TXT_FLAG_SPECIAL = 4, TXT_FLAG_SIMPLE = 8 };
enum { TXT_ERR_CHARS = 1, TXT_ERR_WORD = 2, TXT_ERR_TOKENS = 4, TXT_ERR_ROWS = 8 };

constexpr int TXT_MAX_CHARS = 768;  // bytes of one decoded caption
constexpr int TXT_MAX_WORD = 96;    // byte symbols of one pre-token that needs the BPE loop
constexpr int TXT_MAX_TOK = 96;     // BERT tokens of one caption that survive skip_special_tokens
constexpr int TXT_BODY_MAX = 75;    // CLIP ids between BOS and EOS (truncation to 77)

struct TextVocab {
  const int32_t* tok_off;     // [V+1] byte offsets into tok_bytes / tok_cls ('##' already stripped from pieces)
  const uint8_t* tok_bytes;   // normalised (NFC, lowercase) UTF-8 text of every BERT token
  const uint8_t* tok_cls;     // per byte: TXT_CLS_* of its character | TXT_CHAR_START on the character's first byte
  const uint8_t* tok_flags;   // [V] TXT_FLAG_*
  const int32_t* csr_off;     // [V+1] CLIP ids of each token on its own (valid shortcut when TXT_FLAG_SIMPLE)
  const int32_t* csr_tok;
  const int32_t* byte_sym;    // [512] CLIP id of byte b inside a word [b] / as the last byte of a word [256 + b]
  const uint64_t* merge_keys; // hash table: (a << 32 | b) + 1, 0 = empty slot
  const uint32_t* merge_vals; // rank << 16 | merged id
  int merge_bits;             // table size = 1 << merge_bits
  int V;
};

CZ_HD uint32_t txt_merge_lookup(const TextVocab& v, int32_t a, int32_t b) {  // 0xffffffff = no such merge
  const uint64_t key = ((static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b)) + 1;
  const uint32_t mask = (1u << v.merge_bits) - 1u;
  uint32_t slot = static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> (64 - v.merge_bits)) & mask;
  for (;;) {
    const uint64_t k = v.merge_keys[slot];
    if (k == key) return v.merge_vals[slot];
    if (k == 0) return 0xffffffffu;
    slot = (slot + 1) & mask;
  }
}

// Byte-level BPE of one pre-token buf[s, e): appends its CLIP ids to out (at most max_out in total), returns the new
// count; *err |= TXT_ERR_WORD when the word has more byte symbols than TXT_MAX_WORD.
CZ_HD int txt_bpe_word(const TextVocab& v, const uint8_t* buf, int s, int e, int32_t* out, int n_out, int max_out, int* err) {
  int32_t sym[TXT_MAX_WORD];
  int n = e - s;
  if (n > TXT_MAX_WORD) { *err |= TXT_ERR_WORD; n = TXT_MAX_WORD; }
  for (int i = 0; i < n; ++i) sym[i] = v.byte_sym[(i == n - 1 ? 256 : 0) + buf[s + i]];
  while (n > 1) {
    uint32_t best = 0xffffffffu;
    for (int i = 0; i + 1 < n; ++i) {
      const uint32_t r = txt_merge_lookup(v, sym[i], sym[i + 1]);
      if (r != 0xffffffffu && (r >> 16) < (best >> 16)) best = r;  // leftmost on equal rank: strict <
    }
    if (best == 0xffffffffu) break;
    // merge every occurrence of the best pair, left to right
    const uint32_t rank = best >> 16;
    int w = 0;
    for (int i = 0; i < n; ++i) {
      if (i + 1 < n) {
        const uint32_t r = txt_merge_lookup(v, sym[i], sym[i + 1]);
        if (r != 0xffffffffu && (r >> 16) == rank) {
          sym[w++] = static_cast<int32_t>(r & 0xffffu);
          ++i;
          continue;
        }
      }
      sym[w++] = sym[i];
    }
    n = w;
  }
  for (int i = 0; i < n && n_out < max_out; ++i) out[n_out++] = sym[i];
  return n_out;
}

CZ_HD bool txt_is_special(const int* special5, int64_t id) {
  return id == special5[0] || id == special5[1] || id == special5[2] || id == special5[3] || id == special5[4];
}

// One caption: BERT ids row[0..L) with row[pos] replaced by `cand` (pos < 0: no replacement) -> CLIP ids of the
// decoded, normalised, pre-tokenised, BPE-encoded text (no BOS / EOS), at most max_out of them (<= TXT_BODY_MAX is
// what truncation keeps).  Returns the count.
template <typename IdT>
CZ_HD int txt_caption_to_clip(const TextVocab& v, const IdT* row, int L, int pos, int64_t cand, const int* special5,
                              int32_t* out, int max_out, int* err) {
  uint8_t buf[TXT_MAX_CHARS];
  uint8_t cls[TXT_MAX_CHARS];
  int16_t t_start[TXT_MAX_TOK], t_end[TXT_MAX_TOK];
  int32_t t_id[TXT_MAX_TOK];
  int n = 0, nt = 0;
  // ---- WordPiece decode with per-token clean-up
  for (int j = 0; j < L; ++j) {
    const int64_t id = (j == pos) ? cand : static_cast<int64_t>(row[j]);
    if (id < 0 || id >= v.V || txt_is_special(special5, id)) continue;
    const uint8_t fl = v.tok_flags[id];
    if (fl & TXT_FLAG_SPECIAL) continue;
    const int o0 = v.tok_off[id], len = v.tok_off[id + 1] - o0;
    if (nt >= TXT_MAX_TOK) { *err |= TXT_ERR_TOKENS; break; }
    if (n + len + 3 > TXT_MAX_CHARS) { *err |= TXT_ERR_CHARS; break; }
    bool verbatim_piece = false;
    if (nt == 0) {
      if (fl & TXT_FLAG_PIECE) {  // nothing to glue to: the decoder leaves the first token alone, '##' included
        buf[n] = '#'; cls[n++] = TXT_CLS_OTHER | TXT_CHAR_START;
        buf[n] = '#'; cls[n++] = TXT_CLS_OTHER | TXT_CHAR_START;
        verbatim_piece = true;
      }
    } else if (!(fl & (TXT_FLAG_PIECE | TXT_FLAG_GLUE))) {
      buf[n] = ' '; cls[n++] = TXT_CLS_SPACE | TXT_CHAR_START;
    }
    t_start[nt] = static_cast<int16_t>(verbatim_piece ? -1 : n);  // -1: never matches a pre-token start
    for (int i = 0; i < len; ++i) { buf[n] = v.tok_bytes[o0 + i]; cls[n++] = v.tok_cls[o0 + i]; }
    t_end[nt] = static_cast<int16_t>(n);
    t_id[nt++] = static_cast<int32_t>(id);
  }
  // ---- pre-tokenise + BPE
  int n_out = 0, ct = 0, i = 0;
  while (i < n && n_out < max_out) {
    const int c = cls[i] & 3;
    if (c == TXT_CLS_SPACE) { ++i; continue; }
    int e = i;
    if (buf[i] == '\'' && i + 1 < n) {  // 's|'t|'re|'ve|'m|'ll|'d, tried before the character classes
      const uint8_t c1 = buf[i + 1], c2 = i + 2 < n ? buf[i + 2] : 0;
      if (c1 == 's' || c1 == 't' || c1 == 'm' || c1 == 'd') e = i + 2;
      else if ((c1 == 'r' && c2 == 'e') || (c1 == 'v' && c2 == 'e') || (c1 == 'l' && c2 == 'l')) e = i + 3;
    }
    if (e == i) {
      if (c == TXT_CLS_NUMBER) {  // [\p{N}]: one character
        e = i + 1;
        while (e < n && !(cls[e] & TXT_CHAR_START)) ++e;
      } else {                   // [\p{L}]+ or [^\s\p{L}\p{N}]+
        e = i + 1;
        while (e < n && (cls[e] & 3) == c) ++e;
      }
    }
    while (ct < nt && t_end[ct] <= i) ++ct;
    if (ct < nt && t_start[ct] == i && t_end[ct] == e && (v.tok_flags[t_id[ct]] & TXT_FLAG_SIMPLE)) {
      for (int t = v.csr_off[t_id[ct]]; t < v.csr_off[t_id[ct] + 1] && n_out < max_out; ++t) out[n_out++] = v.csr_tok[t];
    } else {
      n_out = txt_bpe_word(v, buf, i, e, out, n_out, max_out, err);
    }
    i = e;
  }
  return n_out;
}

}  // namespace conzic
