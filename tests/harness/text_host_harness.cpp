// CPU build of csrc/text_pipeline.cuh for the test suite: the very code the CUDA kernels compile, exposed through a
// C entry point so tests can check it against the real transformers tokenizer classes on thousands of captions.
// Test infrastructure only (built by tests/test_text_pipeline.py into tests/harness/); never loaded by the product.
#include "../../conzic_b200/csrc/text_pipeline.cuh"

extern "C" int conzic_text_host_tokenize(const int32_t* tok_off, const uint8_t* tok_bytes, const uint8_t* tok_cls,
                                         const uint8_t* tok_flags, const int32_t* csr_off, const int32_t* csr_tok,
                                         const int32_t* byte_sym, const uint64_t* merge_keys, const uint32_t* merge_vals,
                                         int merge_bits, int V, const int64_t* rows, int n_rows, int L,
                                         const int* special5, int32_t* out, int32_t* out_len) {
  conzic::TextVocab v{tok_off, tok_bytes, tok_cls, tok_flags, csr_off, csr_tok, byte_sym, merge_keys, merge_vals, merge_bits, V};
  int err = 0;
  for (int r = 0; r < n_rows; ++r)
    out_len[r] = conzic::txt_caption_to_clip(v, rows + static_cast<size_t>(r) * L, L, -1, 0, special5,
                                             out + static_cast<size_t>(r) * conzic::TXT_BODY_MAX, conzic::TXT_BODY_MAX, &err);
  return err;
}
