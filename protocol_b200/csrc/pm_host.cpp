// pm_host.cpp — host-side pieces of the C ABI that need no GPU:
//   * model-string interning + the (pattern x model) acceptance bit table,
//     hoisting GpuSpecs::meets' substring clause (reference
//     crates/shared/src/models/node.rs:463-484) out of the per-pair loop;
//   * ComputeRequirements::from_str (node.rs:180-374) straight into the ask /
//     option table rows of include/prime_match.h;
//   * the configuration priority sort (crates/orchestrator/src/plugins/
//     node_groups/mod.rs:150-164).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/prime_match.h"
#include "pm_guard.hpp"

namespace {

// Normal form used by node.rs:465,470: ASCII-lowercase, ' ' -> '_'.
std::string normalise(std::string_view s) {
  std::string out(s);
  for (char& c : out) {
    if (c >= 'A' && c <= 'Z') c = char(c - 'A' + 'a');
    else if (c == ' ') c = '_';
  }
  return out;
}
std::string strip_underscores(const std::string& s) {
  std::string out;
  out.reserve(s.size());
  for (char c : s)
    if (c != '_') out.push_back(c);
  return out;
}
std::string_view trim_view(std::string_view s) {
  auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
  while (!s.empty() && ws((unsigned char)s.front())) s.remove_prefix(1);
  while (!s.empty() && ws((unsigned char)s.back())) s.remove_suffix(1);
  return s;
}
inline bool has(const std::string& hay, const std::string& needle) {
  return hay.find(needle) != std::string::npos;
}

struct ModelForm {
  std::string norm, bare;  // with and without underscores
};
struct PatternForm {
  std::vector<ModelForm> alts;  // comma-separated alternatives
};

bool accepts(const PatternForm& p, const ModelForm& m) {
  // node.rs:471-480: any alternative where either string contains the other,
  // compared both with and without underscores.
  for (const ModelForm& alt : p.alts) {
    if (has(m.norm, alt.norm) || has(alt.norm, m.norm) || has(m.bare, alt.bare) ||
        has(alt.bare, m.bare))
      return true;
  }
  return false;
}

}  // namespace

struct pm_interner {
  std::unordered_map<std::string, uint32_t> model_ids, pattern_ids;
  std::vector<ModelForm> models;
  std::vector<PatternForm> patterns;
  std::vector<uint32_t> bits;
  uint32_t words = 0;
  bool dirty = true;
};

extern "C" {

uint32_t pm_abi_version(void) { return PM_ABI_VERSION; }

pm_interner* pm_interner_create(void) { return new (std::nothrow) pm_interner; }
void pm_interner_destroy(pm_interner* in) { delete in; }

uint32_t pm_intern_model(pm_interner* in, const char* spec_model) try {
  if (!in || !spec_model) return PM_NONE;
  std::string key(spec_model);
  auto it = in->model_ids.find(key);
  if (it != in->model_ids.end()) return it->second;
  ModelForm f;
  f.norm = normalise(key);
  f.bare = strip_underscores(f.norm);
  uint32_t id = uint32_t(in->models.size());
  in->models.push_back(std::move(f));
  in->model_ids.emplace(std::move(key), id);
  in->dirty = true;
  return id;
} catch (...) { return PM_NONE; }

uint32_t pm_intern_pattern(pm_interner* in, const char* req_model) try {
  if (!in || !req_model) return PM_NONE;
  std::string key(req_model);
  auto it = in->pattern_ids.find(key);
  if (it != in->pattern_ids.end()) return it->second;
  PatternForm p;
  std::string_view rest(key);
  for (;;) {  // str::split(',') keeps empty pieces
    size_t comma = rest.find(',');
    std::string_view piece = comma == std::string_view::npos ? rest : rest.substr(0, comma);
    ModelForm f;
    f.norm = normalise(trim_view(piece));
    f.bare = strip_underscores(f.norm);
    p.alts.push_back(std::move(f));
    if (comma == std::string_view::npos) break;
    rest.remove_prefix(comma + 1);
  }
  uint32_t id = uint32_t(in->patterns.size());
  in->patterns.push_back(std::move(p));
  in->pattern_ids.emplace(std::move(key), id);
  in->dirty = true;
  return id;
} catch (...) { return PM_NONE; }

int pm_interner_table(pm_interner* in, const uint32_t** bits, uint32_t* n_patterns,
                      uint32_t* n_models, uint32_t* words_per_pattern) try {
  if (!in) return PM_E_INVALID;
  if (in->dirty) {
    const uint32_t nm = uint32_t(in->models.size()), np = uint32_t(in->patterns.size());
    in->words = nm ? (nm + 31) / 32 : 1;
    in->bits.assign(size_t(np ? np : 1) * in->words, 0u);
    for (uint32_t p = 0; p < np; ++p)
      for (uint32_t m = 0; m < nm; ++m)
        if (accepts(in->patterns[p], in->models[m]))
          in->bits[size_t(p) * in->words + (m >> 5)] |= 1u << (m & 31);
    in->dirty = false;
  }
  if (bits) *bits = in->bits.data();
  if (n_patterns) *n_patterns = uint32_t(in->patterns.size());
  if (n_models) *n_models = uint32_t(in->models.size());
  if (words_per_pattern) *words_per_pattern = in->words;
  return PM_OK;
} catch (...) { return pm_guard_rc(); }

// ---- ComputeRequirements::from_str (node.rs:180-374) ------------------------
namespace {
bool to_u32(std::string_view s, uint32_t* out) {  // u32::from_str
  if (!s.empty() && s.front() == '+') s.remove_prefix(1);
  if (s.empty()) return false;
  uint64_t v = 0;
  for (char c : s) {
    if (c < '0' || c > '9') return false;
    v = v * 10 + uint64_t(c - '0');
    if (v > 0xFFFFFFFFull) return false;
  }
  *out = uint32_t(v);
  return true;
}
int fail(char* err, size_t n, const std::string& msg) {
  if (err && n) std::snprintf(err, n, "%s", msg.c_str());
  return PM_E_PARSE;
}
}  // namespace

int pm_parse_requirements(const char* s, pm_interner* interner, pm_ask* ask, pm_gpu_opt* opts,
                          uint32_t max_opts, uint32_t* n_opts_out, char* err, size_t err_len) try {
  if (!s || !ask) return PM_E_INVALID;
  std::vector<pm_gpu_opt> done;
  pm_gpu_opt cur{};
  bool started = false;
  uint32_t flags = PM_A_HAS_REQ, cores = 0, ram = 0, storage = 0;

  std::string_view rest(s);
  bool more = true;
  while (more) {
    size_t semi = rest.find(';');
    std::string_view part = trim_view(semi == std::string_view::npos ? rest : rest.substr(0, semi));
    if (semi == std::string_view::npos) more = false;
    else rest.remove_prefix(semi + 1);
    if (part.empty()) continue;

    size_t eq = part.find('=');
    if (eq == std::string_view::npos)
      return fail(err, err_len, "Invalid key-value pair format: '" + std::string(part) + "'");
    std::string_view key = trim_view(part.substr(0, eq));
    std::string_view val = trim_view(part.substr(eq + 1));
    uint32_t v = 0;
    const bool numeric = to_u32(val, &v);
    auto bad_num = [&]() {
      return fail(err, err_len,
                  "Invalid " + std::string(key) + " value '" + std::string(val) + "'");
    };
    const char* both = "Cannot specify both exact memory and min/max memory";

    if (key == "gpu:count") {
      if (started && (cur.present & PM_O_COUNT)) {  // a second count opens a new OR-option
        done.push_back(cur);
        cur = pm_gpu_opt{};
      }
      started = true;
      if (!numeric) return bad_num();
      cur.present |= PM_O_COUNT;
      cur.count = v;
    } else if (key == "gpu:model") {
      started = true;
      cur.present |= PM_O_MODEL;
      cur.pattern_id = interner ? pm_intern_pattern(interner, std::string(val).c_str()) : 0;
    } else if (key == "gpu:memory_mb") {
      started = true;
      if (cur.present & (PM_O_MEM_MIN | PM_O_MEM_MAX)) return fail(err, err_len, both);
      if (!numeric) return bad_num();
      cur.present |= PM_O_MEM;
      cur.memory_mb = v;
    } else if (key == "gpu:memory_mb_min") {
      started = true;
      if (cur.present & PM_O_MEM) return fail(err, err_len, both);
      if (!numeric) return bad_num();  // reference would panic on unwrap() if max was set
      if ((cur.present & PM_O_MEM_MAX) && cur.memory_mb_max < v)
        return fail(err, err_len, "Invalid gpu:memory_mb_min value '" + std::string(val) +
                                      "': min value is greater than max value");
      cur.present |= PM_O_MEM_MIN;
      cur.memory_mb_min = v;
    } else if (key == "gpu:memory_mb_max") {
      started = true;
      if (cur.present & PM_O_MEM) return fail(err, err_len, both);
      if (!numeric) return bad_num();
      if ((cur.present & PM_O_MEM_MIN) && cur.memory_mb_min > v)
        return fail(err, err_len, "Invalid gpu:memory_mb_max value '" + std::string(val) +
                                      "': max value is less than min value");
      cur.present |= PM_O_MEM_MAX;
      cur.memory_mb_max = v;
    } else if (key == "gpu:total_memory_min") {
      started = true;
      if (!numeric) return bad_num();
      if ((cur.present & PM_O_TOT_MAX) && cur.total_memory_max < v)
        return fail(err, err_len, "Invalid gpu:total_memory_min value '" + std::string(val) +
                                      "': min value is greater than max value");
      cur.present |= PM_O_TOT_MIN;
      cur.total_memory_min = v;
    } else if (key == "gpu:total_memory_max") {
      started = true;
      if (!numeric) return bad_num();
      if ((cur.present & PM_O_TOT_MIN) && cur.total_memory_min > v)
        return fail(err, err_len, "Invalid gpu:total_memory_max value '" + std::string(val) +
                                      "': max value is less than min value");
      cur.present |= PM_O_TOT_MAX;
      cur.total_memory_max = v;
    } else if (key == "cpu:cores") {
      if (!numeric) return bad_num();
      flags |= PM_A_REQ_CPU | PM_A_REQ_CPU_CORES;
      cores = v;
    } else if (key == "ram_mb") {
      if (!numeric) return bad_num();
      flags |= PM_A_REQ_RAM;
      ram = v;
    } else if (key == "storage_gb") {
      if (!numeric) return bad_num();
      flags |= PM_A_REQ_STORAGE;
      storage = v;
    } else {
      return fail(err, err_len, "Unknown requirement key: '" + std::string(key) + "'");
    }
  }
  if (started && cur.present != 0) done.push_back(cur);

  if (n_opts_out) *n_opts_out = uint32_t(done.size());
  if (done.size() > max_opts)
    return fail(err, err_len, "too many GPU options for the caller's buffer");
  for (size_t i = 0; i < done.size(); ++i) opts[i] = done[i];
  ask->flags = flags;
  ask->n_opts = uint32_t(done.size());
  ask->cpu_cores = cores;
  ask->ram_mb = ram;
  ask->storage_gb = storage;
  return PM_OK;
} catch (...) { return pm_guard_rc(); }

// mod.rs:150-164
int pm_sort_configs(const uint32_t* min_group_size, const uint8_t* has_requirements, uint32_t n,
                    uint32_t* perm_out) try {
  if ((n && (!min_group_size || !has_requirements)) || !perm_out) return PM_E_INVALID;
  std::vector<uint32_t> p(n);
  for (uint32_t i = 0; i < n; ++i) p[i] = i;
  std::stable_sort(p.begin(), p.end(), [&](uint32_t a, uint32_t b) {
    if (min_group_size[a] != min_group_size[b]) return min_group_size[a] > min_group_size[b];
    return has_requirements[a] && !has_requirements[b];
  });
  std::copy(p.begin(), p.end(), perm_out);
  return PM_OK;
} catch (...) { return pm_guard_rc(); }

// Keccak-256 (the pre-NIST padding 0x01, as Ethereum uses it) of a short message: one 136-byte block is enough for the
// 40 hex digits EIP-55 hashes.
static void keccak256_short(const unsigned char* msg, size_t len, unsigned char out[32]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
      0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
      0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
      0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
  static const int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
  unsigned char block[136] = {0};
  std::memcpy(block, msg, len);   // len < 136
  block[len] ^= 0x01;
  block[135] ^= 0x80;
  uint64_t st[25] = {0};
  for (int i = 0; i < 17; ++i) {
    uint64_t lane = 0;
    for (int b = 7; b >= 0; --b) lane = (lane << 8) | block[i * 8 + b];
    st[i] = lane;
  }
  auto rotl = [](uint64_t x, int n) { return (x << n) | (x >> (64 - n)); };
  for (int round = 0; round < 24; ++round) {
    uint64_t bc[5];
    for (int i = 0; i < 5; ++i) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; ++i) {
      const uint64_t t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1);
      for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
    }
    uint64_t t = st[1];
    for (int i = 0; i < 24; ++i) {
      const int j = PIL[i];
      const uint64_t keep = st[j];
      st[j] = rotl(t, ROT[i]);
      t = keep;
    }
    for (int j = 0; j < 25; j += 5) {
      uint64_t row[5];
      for (int i = 0; i < 5; ++i) row[i] = st[j + i];
      for (int i = 0; i < 5; ++i) st[j + i] ^= (~row[(i + 1) % 5]) & row[(i + 2) % 5];
    }
    st[0] ^= RC[round];
  }
  for (int i = 0; i < 4; ++i)
    for (int b = 0; b < 8; ++b) out[i * 8 + b] = (unsigned char)(st[i] >> (8 * b));
}

// Address::from_str + Address::to_string of the reference (alloy-primitives 1.1.0): 40 hex digits, optional "0x", any
// case in; "0x" + EIP-55 checksum case out (a hex letter is upper case iff the matching nibble of keccak256(lower-case
// digits) is >= 8).
int pm_address_canonical(const char* address, char out[43]) {
  if (!address || !out) return PM_E_INVALID;
  const char* h = address;
  if (h[0] == '0' && (h[1] == 'x' || h[1] == 'X')) h += 2;
  unsigned char lower[40];
  for (int i = 0; i < 40; ++i) {
    const char c = h[i];
    if (c >= '0' && c <= '9') lower[i] = (unsigned char)c;
    else if (c >= 'a' && c <= 'f') lower[i] = (unsigned char)c;
    else if (c >= 'A' && c <= 'F') lower[i] = (unsigned char)(c - 'A' + 'a');
    else return PM_E_INVALID;   // also a string that ends early
  }
  if (h[40] != '\0') return PM_E_INVALID;
  unsigned char hash[32];
  keccak256_short(lower, 40, hash);
  out[0] = '0';
  out[1] = 'x';
  for (int i = 0; i < 40; ++i) {
    const unsigned nib = (i & 1) ? (hash[i / 2] & 0x0Fu) : (hash[i / 2] >> 4);
    const char c = (char)lower[i];
    out[2 + i] = (c >= 'a' && nib >= 8u) ? (char)(c - 'a' + 'A') : c;
  }
  out[42] = '\0';
  return PM_OK;
}

}  // extern "C"
