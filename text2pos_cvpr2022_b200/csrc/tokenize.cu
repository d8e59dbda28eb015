// Host-side tokeniser of the text encoder (no device code): the rules of LanguageEncoder.forward,
// models/modules.py:60-72 of the reference -- remove '.' and ',', lower-case, split on whitespace, map words through
// the vocabulary (out-of-vocabulary -> 0), zero-pad to the longest description of the batch.  Writes straight into
// caller-provided (pinned) int32 buffers so that the H2D copy needs no intermediate numpy array.
#include <string.h>

#include <string>
#include <unordered_map>

#include "common.cuh"

struct t2p_vocab {
  std::unordered_map<std::string, int32_t> map;
};

namespace {
// str.split() separators within ASCII: space, \t \n \v \f \r and the information separators 0x1c-0x1f
inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }
}  // namespace

extern "C" {

int t2p_vocab_create(const char* const* words, const int32_t* ids, int n, t2p_vocab** out) {
  T2P_REQUIRE(out != nullptr && n >= 0 && (n == 0 || (words != nullptr && ids != nullptr)), T2P_ERR_INVALID, "vocab_create: null argument");
  t2p_vocab* v = new t2p_vocab();
  v->map.reserve((size_t)n * 2 + 1);
  for (int i = 0; i < n; ++i) {
    if (words[i] == nullptr) {
      delete v;
      t2p::set_error("vocab_create: word %d is null", i);
      return T2P_ERR_INVALID;
    }
    v->map[std::string(words[i])] = ids[i];
  }
  *out = v;
  return T2P_OK;
}

int t2p_vocab_destroy(t2p_vocab* v) {
  delete v;
  return T2P_OK;
}

// texts: n_texts strings stored back to back, each terminated by '\0' (total_bytes including the terminators).
// h_tokens [n_texts, max_tokens] (fully written: zero padded), h_lengths [n_texts], *out_max_len = longest row.
// Bytes >= 0x80 are treated as word characters (callers route non-ASCII strings through the Python tokeniser,
// whose lower()/split() are Unicode-aware).  A description longer than max_tokens is an error.
int t2p_tokenize(const t2p_vocab* v, const char* texts, size_t total_bytes, int n_texts, int max_tokens, int32_t* h_tokens,
                 int32_t* h_lengths, int32_t* out_max_len) {
  T2P_REQUIRE(v && texts && h_tokens && h_lengths && n_texts >= 0 && max_tokens >= 1, T2P_ERR_INVALID, "tokenize: bad argument");
  const char* p = texts;
  const char* end = texts + total_bytes;
  int32_t longest = 0;
  std::string word;
  word.reserve(64);
  for (int i = 0; i < n_texts; ++i) {
    T2P_REQUIRE(p < end, T2P_ERR_INVALID, "tokenize: %d texts announced but the buffer holds only %d", n_texts, i);
    int32_t* row = h_tokens + (size_t)i * max_tokens;
    int32_t n = 0;
    word.clear();
    auto flush = [&]() -> bool {
      if (word.empty()) return true;
      if (n >= max_tokens) return false;
      auto it = v->map.find(word);
      row[n++] = it == v->map.end() ? 0 : it->second;
      word.clear();
      return true;
    };
    for (; p < end && *p != '\0'; ++p) {
      const unsigned char c = (unsigned char)*p;
      if (c == '.' || c == ',') continue;
      if (is_space(c)) {
        T2P_REQUIRE(flush(), T2P_ERR_UNSUPPORTED, "tokenize: description %d has more than %d tokens", i, max_tokens);
      } else {
        word.push_back((c >= 'A' && c <= 'Z') ? (char)(c + 32) : (char)c);
      }
    }
    T2P_REQUIRE(flush(), T2P_ERR_UNSUPPORTED, "tokenize: description %d has more than %d tokens", i, max_tokens);
    T2P_REQUIRE(p < end, T2P_ERR_INVALID, "tokenize: text %d is not NUL-terminated inside the buffer", i);
    ++p;  // skip the terminator
    for (int32_t t = n; t < max_tokens; ++t) row[t] = 0;
    h_lengths[i] = n;
    if (n > longest) longest = n;
  }
  if (out_max_len) *out_max_len = longest;
  return T2P_OK;
}

}  // extern "C"
