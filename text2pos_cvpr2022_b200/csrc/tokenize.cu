// Host-side tokeniser of the text encoder (no device code): the rules of LanguageEncoder.forward,
// models/modules.py:60-72 of the reference -- remove '.' and ',', lower-case, split on whitespace, map words through
// the vocabulary (out-of-vocabulary -> 0), zero-pad to the longest description of the batch.  Writes straight into
// caller-provided (pinned) int32 buffers so that the H2D copy needs no intermediate numpy array.
//
// Two implementations of the same rules: t2p_tokenize (host, writes token ids) and t2p_tokenize_device (one CTA per
// description on the GPU: the serving engine ships the raw bytes of a batch -- about as many as the token ids -- and the
// tokens never exist on the host; ~15 KB of text took the host longer than the whole GPU step).
#include <string.h>

#include <mutex>
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
  // device mirror (t2p_vocab_to_device): [Entry table[mask+1] | chars], on device `dev`
  void* d_table = nullptr;
  int dev = -1;
  std::mutex mu;

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


// ---- device tokeniser ----------------------------------------------------------------------------------------------
namespace t2p {
constexpr int TOKD_THREADS = 128;
constexpr int TOKD_MAX_BYTES = 8192;  // longest description the device path takes (the templates give ~250 bytes)

__device__ __forceinline__ int tokd_class(unsigned c) {
  return (c == ' ' || (c >= 9u && c <= 13u) || (c >= 0x1cu && c <= 0x1fu)) ? 1 : (c == '.' || c == ',') ? 2 : 0;
}
// exclusive block scan of one int per thread (TOKD_THREADS = 4 warps); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int tokd_block_exscan(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();  // warp_tot may still be read from the previous scan
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < TOKD_THREADS / 32; ++w) {
    const int t = warp_tot[w];
    base += (w < warp) ? t : 0;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

// stage: [int32 offsets[n_texts + 1] | padding to 16 bytes | text bytes]; description b = bytes [off[b], off[b+1] - 1)
constexpr int TOKD_TABLE_SMEM = 12288;  // hash table + word characters staged in shared memory when they fit (the usual case)

__global__ void __launch_bounds__(TOKD_THREADS)
tokenize_kernel(const uint8_t* __restrict__ stage, int n_texts, int text_base, const t2p_vocab::Entry* __restrict__ table,
                const char* __restrict__ table_chars, uint32_t mask, uint32_t table_bytes, int max_tokens,
                int32_t* __restrict__ tokens, int32_t* __restrict__ lengths) {
  __shared__ uint8_t comp[TOKD_MAX_BYTES];  // punctuation removed, lower-cased, separators stored as 0
  __shared__ __align__(16) uint8_t tab_s[TOKD_TABLE_SMEM];
  __shared__ int warp_tot[TOKD_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int32_t* off = reinterpret_cast<const int32_t*>(stage);
  const int begin = off[b], L = off[b + 1] - 1 - begin;
  int32_t* row = tokens + (size_t)b * max_tokens;
  if (L < 0 || L > TOKD_MAX_BYTES) {  // malformed / too long for this path: flagged, the host raises
    for (int t = tid; t < max_tokens; t += TOKD_THREADS) row[t] = 0;
    if (tid == 0) lengths[b] = -1;
    return;
  }
  // the vocabulary (entries + characters, contiguous on the device) goes to shared memory in the same round trip as the text
  const bool tab_in_smem = table_bytes <= TOKD_TABLE_SMEM;
  if (tab_in_smem) {
    const uint4* src = reinterpret_cast<const uint4*>(table);
    for (uint32_t t = tid; t < (table_bytes + 15) / 16; t += TOKD_THREADS) reinterpret_cast<uint4*>(tab_s)[t] = __ldg(src + t);
  }
  const t2p_vocab::Entry* tab = tab_in_smem ? reinterpret_cast<const t2p_vocab::Entry*>(tab_s) : table;
  const char* tchars = tab_in_smem ? reinterpret_cast<const char*>(tab_s) + (size_t)(mask + 1) * sizeof(t2p_vocab::Entry) : table_chars;
  const uint8_t* text = stage + text_base + begin;
  // 1. drop '.' and ',', lower-case: each thread owns a contiguous segment
  const int seg = (L + TOKD_THREADS - 1) / TOKD_THREADS;
  const int s0 = min(tid * seg, L), s1 = min(s0 + seg, L);
  int keep = 0;
  for (int i = s0; i < s1; ++i) keep += tokd_class(text[i]) != 2;
  int Lc;
  int pos = tokd_block_exscan(keep, warp_tot, &Lc);
  for (int i = s0; i < s1; ++i) {
    const unsigned c = text[i];
    const int cls = tokd_class(c);
    if (cls == 2) continue;
    comp[pos++] = cls == 1 ? (uint8_t)0 : (uint8_t)((c >= 'A' && c <= 'Z') ? c + 32 : c);
  }
  __syncthreads();
  // 2. word starts -> word index; 3. the thread that owns a start hashes the word and looks it up
  const int segc = (Lc + TOKD_THREADS - 1) / TOKD_THREADS;
  const int c0 = min(tid * segc, Lc), c1 = min(c0 + segc, Lc);
  int starts = 0;
  for (int i = c0; i < c1; ++i) starts += (comp[i] != 0 && (i == 0 || comp[i - 1] == 0)) ? 1 : 0;
  int n_words;
  int widx = tokd_block_exscan(starts, warp_tot, &n_words);
  for (int i = c0; i < c1; ++i) {
    if (comp[i] == 0 || (i != 0 && comp[i - 1] != 0)) continue;
    uint64_t h = 1469598103934665603ull;
    int e = i;
    for (; e < Lc && comp[e] != 0; ++e) h = (h ^ comp[e]) * 1099511628211ull;
    const uint32_t len = (uint32_t)(e - i);
    int32_t id = 0;
    for (uint32_t slot = (uint32_t)h & mask;; slot = (slot + 1) & mask) {
      const t2p_vocab::Entry en = tab[slot];
      if (en.len == 0) break;
      if (en.hash == h && en.len == len) {
        bool same = true;
        for (uint32_t j = 0; j < len; ++j) same = same && (uint8_t)tchars[en.off + j] == comp[i + j];
        if (same) {
          id = en.id;
          break;
        }
      }
    }
    if (widx < max_tokens) row[widx] = id;
    ++widx;
  }
  for (int t = n_words + tid; t < max_tokens; t += TOKD_THREADS) row[t] = 0;
  if (tid == 0) lengths[b] = n_words <= max_tokens ? n_words : max_tokens + 1;  // max_tokens + 1: too many tokens (flag)
}
}  // namespace t2p

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
  if (v != nullptr && v->d_table != nullptr) cudaFree(v->d_table);
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

int t2p_vocab_to_device(t2p_vocab* v) {
  T2P_REQUIRE(v != nullptr, T2P_ERR_INVALID, "vocab_to_device: null vocabulary");
  std::lock_guard<std::mutex> lock(v->mu);
  int dev = 0;
  T2P_CUDA(cudaGetDevice(&dev));
  if (v->d_table != nullptr) {
    T2P_REQUIRE(v->dev == dev, T2P_ERR_INVALID, "vocab_to_device: vocabulary already lives on device %d (current %d)", v->dev, dev);
    return T2P_OK;
  }
  const size_t tbytes = v->table.size() * sizeof(t2p_vocab::Entry);
  void* d = nullptr;
  T2P_CUDA(cudaMalloc(&d, tbytes + v->chars.size() + 32));
  cudaError_t e = cudaMemcpy(d, v->table.data(), tbytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !v->chars.empty())
    e = cudaMemcpy(static_cast<char*>(d) + tbytes, v->chars.data(), v->chars.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d);
    t2p::set_error("vocab_to_device: %s", cudaGetErrorString(e));
    return T2P_ERR_CUDA;
  }
  v->d_table = d;
  v->dev = dev;
  return T2P_OK;
}

size_t t2p_stage_texts_capacity(int n_texts, size_t text_bytes) {
  return t2p::align_up((size_t)(n_texts + 1) * sizeof(int32_t), 16) + t2p::align_up(text_bytes, 16);
}

// Host half of the device tokeniser: lays the batch out in the (pinned) staging buffer as
// [int32 byte offsets [n_texts + 1] | padding to 16 bytes | the NUL-terminated strings back to back].
int t2p_stage_texts(const char* texts, size_t total_bytes, int n_texts, void* h_stage, size_t stage_capacity, void* d_stage,
                    t2p_stream stream, size_t* used_bytes, int* all_ascii) {
  T2P_REQUIRE(texts && h_stage && n_texts >= 0, T2P_ERR_INVALID, "stage_texts: bad argument");
  const size_t head = t2p::align_up((size_t)(n_texts + 1) * sizeof(int32_t), 16);
  T2P_REQUIRE(head + total_bytes <= stage_capacity, T2P_ERR_WORKSPACE, "stage_texts: %zu bytes of text + %zu of offsets exceed the staging buffer (%zu)",
              total_bytes, head, stage_capacity);
  int32_t* off = static_cast<int32_t*>(h_stage);
  const char* p = texts;
  const char* end = texts + total_bytes;
  for (int i = 0; i < n_texts; ++i) {
    T2P_REQUIRE(p < end, T2P_ERR_INVALID, "stage_texts: %d texts announced but the buffer holds only %d", n_texts, i);
    off[i] = (int32_t)(p - texts);
    const void* z = memchr(p, 0, (size_t)(end - p));
    T2P_REQUIRE(z != nullptr, T2P_ERR_INVALID, "stage_texts: text %d is not NUL-terminated inside the buffer", i);
    p = static_cast<const char*>(z) + 1;
  }
  off[n_texts] = (int32_t)(p - texts);
  memcpy(static_cast<char*>(h_stage) + head, texts, total_bytes);
  unsigned char acc = 0;
  for (size_t i = 0; i < total_bytes; ++i) acc |= (unsigned char)texts[i];
  const bool ascii = (acc & 0x80) == 0;
  if (all_ascii) *all_ascii = ascii ? 1 : 0;
  const size_t used = head + t2p::align_up(total_bytes, 16);
  if (used_bytes) *used_bytes = used;
  if (d_stage != nullptr && ascii)
    T2P_CUDA(cudaMemcpyAsync(d_stage, h_stage, used <= stage_capacity ? used : stage_capacity, cudaMemcpyHostToDevice, t2p::as_stream(stream)));
  return T2P_OK;
}

int t2p_tokenize_device(const t2p_vocab* v, const void* d_stage, int n_texts, int max_tokens, int32_t* d_tokens,
                        int32_t* d_lengths, t2p_stream stream) {
  T2P_REQUIRE(v && d_stage && d_tokens && d_lengths && n_texts >= 0 && max_tokens >= 1, T2P_ERR_INVALID, "tokenize_device: bad argument");
  T2P_REQUIRE(v->d_table != nullptr, T2P_ERR_INVALID, "tokenize_device: call t2p_vocab_to_device first");
  if (n_texts == 0) return T2P_OK;
  const int text_base = (int)t2p::align_up((size_t)(n_texts + 1) * sizeof(int32_t), 16);
  const t2p_vocab::Entry* table = static_cast<const t2p_vocab::Entry*>(v->d_table);
  const char* chars = static_cast<const char*>(v->d_table) + v->table.size() * sizeof(t2p_vocab::Entry);
  const uint32_t table_bytes = (uint32_t)(v->table.size() * sizeof(t2p_vocab::Entry) + v->chars.size());
  t2p::tokenize_kernel<<<n_texts, t2p::TOKD_THREADS, 0, t2p::as_stream(stream)>>>(static_cast<const uint8_t*>(d_stage), n_texts, text_base,
                                                                                 table, chars, v->mask, table_bytes, max_tokens, d_tokens,
                                                                                 d_lengths);
  T2P_LAUNCH_CHECK();
  return T2P_OK;
}

}  // extern "C"
