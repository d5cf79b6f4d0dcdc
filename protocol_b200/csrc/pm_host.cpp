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

}  // extern "C"
