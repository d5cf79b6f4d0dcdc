// Host-side rows of SURVEY 8(f), timed through the C ABI (no Python in the loop):
//   (f-1) DiscoveryMonitor ingest  -> pm_plugin_sync_discovery_json / pm_plugin_sync_discovery
//   (f-2) heartbeat fast path      -> pm_scheduler_get_task_for_node
//   the management pass            -> pm_plugin_try_form_new_groups (host snapshot + GPU pass + publish)
// Build:  g++ -O2 -std=c++17 tools/host_bench.cpp -Iinclude -Lprotocol_b200 -lprime_match -lpthread
//             -Wl,-rpath,$PWD/protocol_b200 -o /tmp/pm_host_bench
// Run:    /tmp/pm_host_bench [n_nodes=1000000] [n_tasks=2000] [noconfig]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "prime_match.h"

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static std::string addr_of(uint32_t i) {
  char b[48];
  std::snprintf(b, sizeof b, "0x%040x", i + 1);
  return b;
}
#define CHECK(call)                                                                            \
  do {                                                                                         \
    const int rc_ = (call);                                                                    \
    if (rc_ != PM_OK) {                                                                        \
      std::fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, plugin ? pm_plugin_last_error(plugin) : ""); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

int main(int argc, char** argv) {
  const uint32_t N = argc > 1 ? (uint32_t)std::atoll(argv[1]) : 1000000u;
  const uint32_t T = argc > 2 ? (uint32_t)std::atoll(argv[2]) : 2000u;
  static const char* kModels[4] = {"NVIDIA H100 80GB HBM3", "NVIDIA A100-SXM4-80GB", "NVIDIA GeForce RTX 4090", "NVIDIA L40S"};
  static const uint32_t kMem[4] = {81559, 81920, 24564, 46068};
  pm_plugin* plugin = nullptr;
  pm_engine* engine = nullptr;
  pm_cfg cfg;
  std::memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = pm_abi_version();
  if (pm_create(&cfg, &engine) != PM_OK) {
    std::printf("# no CUDA device: %s -- ingest and the default scheduler only\n", pm_last_error(nullptr));
    engine = nullptr;
  }
  CHECK(pm_plugin_create(engine, nullptr, &plugin));
  const bool no_configs = argc > 3 && std::strcmp(argv[3], "noconfig") == 0;   // Scheduler with the default NewestTaskPlugin
  if (!no_configs) {
    CHECK(pm_plugin_add_config(plugin, "pair-h100", 2, 2, "gpu:count=8;gpu:model=H100"));
    CHECK(pm_plugin_add_config(plugin, "solo", 1, 1, nullptr));
  }
  CHECK(pm_plugin_seal_configs(plugin));
  std::vector<std::string> addrs(N);
  for (uint32_t i = 0; i < N; ++i) addrs[i] = addr_of(i);

  // ---- (f-1) ingest: the JSON body of GET {discovery}/api/pool/{id} --------------------------------
  std::string body = "{\"success\":true,\"data\":[";
  body.reserve((size_t)N * 420);
  for (uint32_t i = 0; i < N; ++i) {
    char b[1024];
    const uint32_t m = i & 3u;
    std::snprintf(b, sizeof b,
                  "%s{\"id\":\"%s\",\"provider_address\":\"%s\",\"ip_address\":\"10.%u.%u.%u\",\"port\":%u,\"compute_pool_id\":1,"
                  "\"compute_specs\":{\"gpu\":{\"count\":8,\"model\":\"%s\",\"memory_mb\":%u,\"indices\":[0,1,2,3,4,5,6,7]},"
                  "\"cpu\":{\"cores\":%u,\"model\":\"x\"},\"ram_mb\":%u,\"storage_gb\":%u,\"storage_path\":\"/data\"},"
                  "\"is_validated\":true,\"is_active\":true,\"is_provider_whitelisted\":true,\"is_blacklisted\":false,"
                  "\"last_updated\":\"2025-05-01T12:00:00.250Z\",\"created_at\":null,"
                  "\"location\":{\"latitude\":%.4f,\"longitude\":%.4f,\"city\":null,\"region\":null,\"country\":\"CA\"},"
                  "\"latest_balance\":\"1000000000000000000\"}",
                  i ? "," : "", addrs[i].c_str(), addrs[i].c_str(), (i >> 16) & 255u, (i >> 8) & 255u, i & 255u, 8000u + (i % 1000u),
                  kModels[m], kMem[m], 32u << (i % 3u), 65536u << (i % 3u), 500u << (i % 4u), 25.0 + (i % 4000) * 0.01,
                  -120.0 + (i % 5000) * 0.01);
    body += b;
  }
  body += "]}";
  uint32_t n_new = 0;
  double t0 = now_s();
  CHECK(pm_plugin_sync_discovery_json(plugin, body.c_str(), body.size(), 1746100000000ll, 1, &n_new));
  double dt = now_s() - t0;
  std::printf("ingest_json_first   nodes=%u new=%u bytes=%zu  %.3f s  %.3g nodes/s  %.3g MB/s\n", N, n_new, body.size(), dt, N / dt, body.size() / dt / 1e6);
  t0 = now_s();
  CHECK(pm_plugin_sync_discovery_json(plugin, body.c_str(), body.size(), 1746100060000ll, 1, &n_new));
  dt = now_s() - t0;
  std::printf("ingest_json_resync  nodes=%u new=%u  %.3f s  %.3g nodes/s\n", N, n_new, dt, N / dt);
  { std::string().swap(body); }

  // ---- the status updater's part (out of scope) stands in as one upsert per node: Healthy + p2p id ----
  t0 = now_s();
  for (uint32_t i = 0; i < N; ++i) {
    pm_node_desc d;
    std::memset(&d, 0, sizeof d);
    const uint32_t m = i & 3u;
    d.address = addrs[i].c_str();
    d.status = 2;   // Healthy
    d.p2p_id = "12D3KooWbench";
    d.spec_flags = PM_W_HAS_SPECS | PM_W_HAS_GPU | PM_W_HAS_GPU_COUNT | PM_W_HAS_GPU_MODEL | PM_W_HAS_GPU_MEM | PM_W_HAS_CPU | PM_W_HAS_CPU_CORES |
                   PM_W_HAS_RAM | PM_W_HAS_STORAGE;
    d.gpu_count = 8; d.gpu_mem_mb = kMem[m]; d.gpu_model = kModels[m];
    d.cpu_cores = 32u << (i % 3u); d.ram_mb = 65536u << (i % 3u); d.storage_gb = 500u << (i % 4u);
    d.has_location = 1; d.lat = 25.0 + (i % 4000) * 0.01; d.lon = -120.0 + (i % 5000) * 0.01;
    CHECK(pm_plugin_upsert_node(plugin, &d));
  }
  dt = now_s() - t0;
  std::printf("upsert_node         nodes=%u  %.3f s  %.3g nodes/s\n", N, dt, N / dt);

  // ---- tasks -----------------------------------------------------------------------------------------
  std::vector<std::string> task_ids(T);
  for (uint32_t i = 0; i < T; ++i) {
    char b[64];
    std::snprintf(b, sizeof b, "00000000-0000-4000-8000-%012x", i);
    task_ids[i] = b;
    pm_task_desc td;
    std::memset(&td, 0, sizeof td);
    td.id = task_ids[i].c_str(); td.name = "bench-task"; td.image = "image"; td.created_at = i;
    const char* topo_a[1] = {"solo"};
    const char* topo_b[1] = {"pair-h100"};
    td.scheduling = 3; td.allowed_topologies = (i & 1u) ? topo_a : topo_b; td.n_allowed_topologies = 1;
    pm_kv env[2] = {{"RANK", "${GROUP_INDEX}"}, {"WORLD_SIZE", "${GROUP_SIZE}"}};
    td.env_vars = env; td.n_env_vars = 2; td.has_env_vars = 1;
    CHECK(pm_plugin_add_task(plugin, &td));
  }

  // ---- the management pass ---------------------------------------------------------------------------
  uint32_t n_formed = 0;
  if (engine && !no_configs) {
    t0 = now_s();
    CHECK(pm_plugin_try_form_new_groups(plugin, &n_formed));
    dt = now_s() - t0;
    std::printf("try_form_new_groups nodes=%u configs=2 groups_formed=%u  %.3f s (host snapshot + GPU pass + publish)\n", N, n_formed, dt);
    t0 = now_s();
    CHECK(pm_plugin_try_form_new_groups(plugin, &n_formed));
    dt = now_s() - t0;
    std::printf("try_form_new_groups (steady state: everybody grouped) groups_formed=%u  %.3f s\n", n_formed, dt);
    // churn: 1000 nodes die (their groups dissolve) and come back healthy; the pass re-forms what it can from row deltas
    for (uint32_t k = 0; k < 1000 && k * 997u < N; ++k) CHECK(pm_plugin_set_node_status(plugin, addrs[k * 997u].c_str(), 4));
    for (uint32_t k = 0; k < 1000 && k * 997u < N; ++k) CHECK(pm_plugin_set_node_status(plugin, addrs[k * 997u].c_str(), 2));
    t0 = now_s();
    CHECK(pm_plugin_try_form_new_groups(plugin, &n_formed));
    dt = now_s() - t0;
    std::printf("try_form_new_groups (after 1000 nodes died and came back) groups_formed=%u  %.3f s\n", n_formed, dt);
    t0 = now_s();
    uint32_t n_merged = 0;
    CHECK(pm_plugin_try_merge_solo_groups(plugin, &n_merged));
    dt = now_s() - t0;
    std::printf("try_merge_solo_groups groups_merged=%u  %.3f s\n", n_merged, dt);
    t0 = now_s();
    CHECK(pm_plugin_try_form_new_groups(plugin, &n_formed));
    dt = now_s() - t0;
    std::printf("try_form_new_groups (steady state again) groups_formed=%u  %.3f s\n", n_formed, dt);
  }

  if (!engine && !no_configs) {   // no GPU here: groups come in the way they do at start-up, from the stored state
    t0 = now_s();
    for (uint32_t i = 0; i < N; ++i) {
      char gid[24];
      std::snprintf(gid, sizeof gid, "%x", i + 1);
      const char* members[1] = {addrs[i].c_str()};
      CHECK(pm_plugin_restore_group(plugin, gid, "solo", members, 1, nullptr, -1));
    }
    dt = now_s() - t0;
    std::printf("restore_group       groups=%u  %.3f s  %.3g groups/s\n", N, dt, N / dt);
  }

  // ---- ingest -> columnar: the SoA snapshot a management pass uploads -------------------------------
  {
    std::vector<pm_worker_a> ta(N);
    std::vector<pm_worker_b> tb(N);
    std::vector<double> tlat(N), tlon(N);
    std::vector<uint32_t> trank(N);
    uint32_t rows = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = now_s();
      CHECK(pm_plugin_export_tables(plugin, ta.data(), tb.data(), tlat.data(), tlon.data(), trank.data(), N, &rows));
      dt = now_s() - t0;
      std::printf("export_tables       rows=%u  %.3f s  %.3g rows/s  (%s)\n", rows, dt, rows / dt,
                  rep ? "address ranks cached" : "first: address ranks sorted");
    }
  }

  // ---- (f-2) heartbeats ------------------------------------------------------------------------------
  const uint32_t H = N < 1000000u ? N : 1000000u;
  for (int threads : {1, 1, 2, 4, 8}) {
    std::atomic<uint64_t> with_task{0};
    std::atomic<int> bad{0};
    t0 = now_s();
    std::vector<std::thread> pool;
    for (int th = 0; th < threads; ++th)
      pool.emplace_back([&, th]() {
        std::vector<char> buf(1 << 14);
        uint64_t mine = 0;
        for (uint32_t i = th; i < H; i += threads) {
          if (pm_scheduler_get_task_for_node(plugin, addrs[i].c_str(), buf.data(), buf.size()) != PM_OK) { bad = 1; break; }
          if (std::strstr(buf.data(), "\"current_task\":null") == nullptr) ++mine;
        }
        with_task += mine;
      });
    for (auto& t : pool) t.join();
    dt = now_s() - t0;
    if (bad) { std::fprintf(stderr, "heartbeat failed: %s\n", pm_plugin_last_error(plugin)); return 1; }
    std::printf("heartbeat           calls=%u threads=%d with_task=%llu tasks=%u  %.3f s  %.3g heartbeats/s\n", H, threads,
                (unsigned long long)with_task.load(), T, dt, H / dt);
  }
  pm_plugin_destroy(plugin);
  if (engine) pm_destroy(engine);
  return 0;
}
