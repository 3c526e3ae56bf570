// test_accounter.cpp — the reference's own Accounter tests (pkg/flow/account_test.go:47-217), restated against
// the C++ host mirror running on the GPU engine.  Built and run by tests/test_gpu_host_cpp.py (needs a B200).
#include <cassert>
#include <cstdio>
#include <map>
#include <memory>
#include <thread>

#include "accounter.hpp"

using namespace flowagg;
using namespace std::chrono_literals;

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static fa_flow_id key(const uint8_t src[4], const uint8_t dst[4], uint16_t sport, uint16_t dport) {
    fa_flow_id k{};
    k.src_ip[10] = k.src_ip[11] = 0xff; memcpy(k.src_ip + 12, src, 4);
    k.dst_ip[10] = k.dst_ip[11] = 0xff; memcpy(k.dst_ip + 12, dst, 4);
    k.src_port = sport; k.dst_port = dport;
    return k;
}
static RawRecord rec(const fa_flow_id& k, uint64_t bytes, uint64_t ts) {
    RawRecord r{};
    r.id = k; r.metrics.bytes = bytes; r.metrics.packets = 1; r.metrics.start_mono_time_ts = ts;
    r.metrics.end_mono_time_ts = ts; r.metrics.flags = 1;
    return r;
}
static std::string kstr(const fa_flow_id& k) { return std::string(reinterpret_cast<const char*>(&k), 39); }

static const uint8_t A1[4] = {0x12, 0x34, 0x56, 0x78}, A2[4] = {0xaa, 0xbb, 0xcc, 0xdd}, D1[4] = {0x43, 0x21, 0x00, 0xff},
                     D2[4] = {0x11, 0x22, 0x33, 0x44};

static int TestEvict_MaxEntries() {                       // account_test.go:47-128
    const fa_flow_id k1 = key(A1, D1, 333, 8080), k2 = key(A2, D1, 12, 8080), k3 = key(A1, D2, 333, 443);
    const uint64_t now = 1661272402ull * 1000000000ull;
    Metrics m;
    auto acc = NewAccounter(2, std::chrono::hours(1), [&] { return now; }, [] { return uint64_t(1000); }, &m);
    Chan<RawRecord> inputs(20); Chan<std::vector<Record>> evictor(20);
    std::thread t([&] { acc->Account(inputs, evictor); });
    inputs.send(rec(k1, 123, 123)); inputs.send(rec(k2, 456, 456)); inputs.send(rec(k1, 321, 789));
    std::this_thread::sleep_for(20ms);
    CHECK(evictor.len() == 0);                            // requireNoEviction
    inputs.send(rec(k3, 111, 888));
    inputs.close();                                       // flushes the batch: the 4th record triggers the "full" eviction
    t.join();
    auto r = evictor.try_recv();
    CHECK(r && r->size() == 2);
    std::map<std::string, Record> got;
    for (auto& x : *r) got[kstr(x.ID)] = x;
    const Record& g1 = got.at(kstr(k1)); const Record& g2 = got.at(kstr(k2));
    CHECK(g1.Metrics.bytes == 444 && g1.Metrics.packets == 2 && g1.Metrics.start_mono_time_ts == 123 &&
          g1.Metrics.end_mono_time_ts == 789 && g1.Metrics.flags == 1);
    CHECK(g1.TimeFlowStart == now - (1000 - 123) && g1.TimeFlowEnd == now - (1000 - 789));
    CHECK(g2.Metrics.bytes == 456 && g2.Metrics.packets == 1 && g2.Metrics.start_mono_time_ts == 456 &&
          g2.Metrics.end_mono_time_ts == 456);
    CHECK(g2.TimeFlowStart == now - (1000 - 456) && g2.TimeFlowEnd == now - (1000 - 456));
    auto last = evictor.try_recv();                       // closing eviction holds k3 only
    CHECK(last && last->size() == 1 && kstr((*last)[0].ID) == kstr(k3) && (*last)[0].Metrics.bytes == 111);
    CHECK(m.evictions_full == 1 && m.evictions_closing == 1 && m.evicted_flows == 3);
    return 0;
}

static int TestEvict_Period() {                           // account_test.go:130-217
    const fa_flow_id k1 = key(A1, D1, 333, 8080);
    const uint64_t now = 1661272402ull * 1000000000ull;
    Metrics m;
    auto acc = NewAccounter(200, 20ms, [&] { return now; }, [] { return uint64_t(1000); }, &m);
    Chan<RawRecord> inputs(20); Chan<std::vector<Record>> evictor(20);
    std::thread t([&] { acc->Account(inputs, evictor); });
    inputs.send(rec(k1, 10, 123)); inputs.send(rec(k1, 10, 456)); inputs.send(rec(k1, 10, 789));
    std::this_thread::sleep_for(60ms);                    // forcing at least one eviction here
    inputs.send(rec(k1, 10, 1123)); inputs.send(rec(k1, 10, 1456));
    std::this_thread::sleep_for(60ms);
    auto a = evictor.try_recv(); auto b = evictor.try_recv();
    CHECK(a && a->size() == 1 && b && b->size() == 1);
    CHECK((*a)[0].Metrics.bytes == 30 && (*a)[0].Metrics.packets == 3 && (*a)[0].Metrics.start_mono_time_ts == 123 &&
          (*a)[0].Metrics.end_mono_time_ts == 789 && (*a)[0].Metrics.flags == 1);
    CHECK((*a)[0].TimeFlowStart == now - 1000 + 123 && (*a)[0].TimeFlowEnd == now - 1000 + 789);
    CHECK((*b)[0].Metrics.bytes == 20 && (*b)[0].Metrics.packets == 2 && (*b)[0].Metrics.start_mono_time_ts == 1123 &&
          (*b)[0].Metrics.end_mono_time_ts == 1456);
    CHECK((*b)[0].TimeFlowStart == now - 1000 + 1123 && (*b)[0].TimeFlowEnd == now - 1000 + 1456);
    std::this_thread::sleep_for(60ms);
    CHECK(evictor.len() == 0);                            // no more flows are evicted
    inputs.close(); t.join();
    return 0;
}

int main() {
    if (int rc = TestEvict_MaxEntries()) return rc;
    printf("ok TestEvict_MaxEntries\n");
    if (int rc = TestEvict_Period()) return rc;
    printf("ok TestEvict_Period\n");
    return 0;
}
