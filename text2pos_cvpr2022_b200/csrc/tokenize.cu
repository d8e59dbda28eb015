// Host-side tokeniser of the text encoder (no device code): the rules of LanguageEncoder.forward,
// models/modules.py:60-72 of the reference -- remove '.' and ',', lower-case, split on whitespace, map words through
// the vocabulary (out-of-vocabulary -> 0), zero-pad to the longest description of the batch.  Writes straight into
// caller-provided (pinned) int32 buffers so that the H2D copy needs no intermediate numpy array.
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

// Open-addressing table keyed by the FNV-1a hash of the lower-cased word: the scan hashes while it lower-cases, so a
// lookup is one probe + one memcmp (the vocabulary of the hint templates has ~50 words).
struct t2p_vocab {
  struct Entry {
    uint64_t hash;
    int32_t id;
    uint32_t len;   // 0 = empty slot
    uint32_t off;   // into `chars`
  };
  std::vector<Entry> table;
  std::string chars;
  uint32_t mask = 0;

  void insert(const char* w, uint32_t len, uint64_t h, int32_t id) {
    for (uint32_t i = (uint32_t)h & mask;; i = (i + 1) & mask) {
      Entry& e = table[i];
      if (e.len == 0) {
        e.hash = h; e.id = id; e.len = len; e.off = (uint32_t)chars.size();
        chars.append(w, len);
        return;
      }
      if (e.hash == h && e.len == len && memcmp(chars.data() + e.off, w, len) == 0) {
        e.id = id;  // later duplicates win, like dict assignment
        return;
      }
    }
  }
  inline int32_t find(const char* w, uint32_t len, uint64_t h) const {
    for (uint32_t i = (uint32_t)h & mask;; i = (i + 1) & mask) {
      const Entry& e = table[i];
      if (e.len == 0) return 0;  // out of vocabulary
      if (e.hash == h && e.len == len && memcmp(chars.data() + e.off, w, len) == 0) return e.id;
    }
  }
};

namespace {
constexpr uint64_t kFnvBasis = 1469598103934665603ull, kFnvPrime = 1099511628211ull;
inline uint64_t fnv1a(const char* w, size_t len) {
  uint64_t h = kFnvBasis;
  for (size_t i = 0; i < len; ++i) h = (h ^ (unsigned char)w[i]) * kFnvPrime;
  return h;
}
// character classes: 0 = word character, 1 = separator (str.split(): space, \t \n \v \f \r, 0x1c-0x1f), 2 = removed ('.' ',')
struct CharClass {
  unsigned char cls[256];
  unsigned char lower[256];
  CharClass() {
    for (int c = 0; c < 256; ++c) {
      cls[c] = (c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f)) ? 1 : (c == '.' || c == ',') ? 2 : 0;
      lower[c] = (unsigned char)((c >= 'A' && c <= 'Z') ? c + 32 : c);
    }
  }
};
const CharClass kChars;
}  // namespace

extern "C" {

int t2p_vocab_create(const char* const* words, const int32_t* ids, int n, t2p_vocab** out) {
  T2P_REQUIRE(out != nullptr && n >= 0 && (n == 0 || (words != nullptr && ids != nullptr)), T2P_ERR_INVALID, "vocab_create: null argument");
  t2p_vocab* v = new t2p_vocab();
  uint32_t cap = 16;
  while (cap < (uint32_t)n * 4u) cap <<= 1;
  v->table.assign(cap, t2p_vocab::Entry{0, 0, 0, 0});
  v->mask = cap - 1;
  for (int i = 0; i < n; ++i) {
    if (words[i] == nullptr) {
      delete v;
      t2p::set_error("vocab_create: word %d is null", i);
      return T2P_ERR_INVALID;
    }
    const size_t len = strlen(words[i]);
    if (len == 0) continue;  // split() never yields an empty word
    v->insert(words[i], (uint32_t)len, fnv1a(words[i], len), ids[i]);
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
  std::string big;  // words longer than the stack buffer (never in the vocabulary of the templates, but legal input)
  for (int i = 0; i < n_texts; ++i) {
    T2P_REQUIRE(p < end, T2P_ERR_INVALID, "tokenize: %d texts announced but the buffer holds only %d", n_texts, i);
    int32_t* row = h_tokens + (size_t)i * max_tokens;
    int32_t n = 0;
    char word[64];
    uint32_t wl = 0;
    uint64_t h = kFnvBasis;
    big.clear();
    for (;; ++p) {
      T2P_REQUIRE(p < end, T2P_ERR_INVALID, "tokenize: text %d is not NUL-terminated inside the buffer", i);
      const unsigned char c = (unsigned char)*p;
      const unsigned char cls = c == 0 ? 1 : kChars.cls[c];
      if (cls == 0) {
        const unsigned char lc = kChars.lower[c];
        if (wl < sizeof(word)) word[wl] = (char)lc;
        else {
          if (wl == sizeof(word)) big.assign(word, sizeof(word));
          big.push_back((char)lc);
        }
        ++wl;
        h = (h ^ lc) * kFnvPrime;
      } else if (cls == 1) {
        if (wl != 0) {
          T2P_REQUIRE(n < max_tokens, T2P_ERR_UNSUPPORTED, "tokenize: description %d has more than %d tokens", i, max_tokens);
          row[n++] = v->find(wl <= sizeof(word) ? word : big.data(), wl, h);
          wl = 0;
          h = kFnvBasis;
          big.clear();
        }
        if (c == 0) break;
      }
    }
    ++p;  // skip the terminator
    for (int32_t t = n; t < max_tokens; ++t) row[t] = 0;
    h_lengths[i] = n;
    if (n > longest) longest = n;
  }
  if (out_max_len) *out_max_len = longest;
  return T2P_OK;
}

}  // extern "C"
