// accounter.hpp — compiled-language host mirror of the reference's pipeline stage for this path, on top of the
// C ABI (include/flowagg.h).  The reference host is Go; the Go toolchain is absent from the build image, so the
// stage is mirrored in C++ with the same names, argument meaning and behaviour:
//
//   flowagg::Accounter            <->  flow.Accounter            (pkg/flow/account.go:19-124)
//   flowagg::NewAccounter(...)    <->  flow.NewAccounter(maxEntries, evictTimeout, clock, monoClock, ...) (:34-53)
//   Accounter::Account(in, out)   <->  (*Accounter).Account(in <-chan *RawRecord, out chan<- []*Record)   (:58-100)
//   flowagg::NewRecord            <->  model.NewRecord time conversion (pkg/model/record.go:82-97)
//   flowagg::Chan<T>              <->  a Go channel (buffered, closable)
//
// The GPU engine wants batches, so Account() buffers RawRecords and hands them to fa_ingest when the buffer is
// full, when the eviction ticker fires and when the input closes; the "cache is full" rule (account.go:85-94)
// is honoured exactly through FA_FULL.
#pragma once

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/flowagg.h"

namespace flowagg {

using RawRecord = fa_flow_record;                       // model.RawRecord (record.go:63)

struct Record {                                         // model.Record, the fields this path fills (record.go:66-80)
    fa_flow_id      ID;
    fa_flow_metrics Metrics;
    uint64_t        TimeFlowStart;                      // unix ns
    uint64_t        TimeFlowEnd;
};

// NewRecord's time arithmetic: now - (monotonicNow - mono), all uint64 (record.go:90-97).
inline Record NewRecord(const fa_flow_id& key, const fa_flow_metrics& m, uint64_t now_unix_ns, uint64_t mono_now_ns) {
    Record r;
    r.ID = key; r.Metrics = m;
    r.TimeFlowStart = now_unix_ns - (mono_now_ns - m.start_mono_time_ts);
    r.TimeFlowEnd = now_unix_ns - (mono_now_ns - m.end_mono_time_ts);
    return r;
}

template <typename T>
class Chan {                                            // a buffered Go channel
public:
    explicit Chan(size_t cap) : cap_(cap) {}
    void send(T v) {
        std::unique_lock<std::mutex> lk(mu_);
        not_full_.wait(lk, [&] { return q_.size() < cap_ || closed_; });
        if (closed_) throw std::runtime_error("send on closed channel");
        q_.push_back(std::move(v));
        not_empty_.notify_one();
    }
    void close() { std::lock_guard<std::mutex> lk(mu_); closed_ = true; not_empty_.notify_all(); not_full_.notify_all(); }
    // recv with deadline: value, or nullopt with *closed = true when the channel is closed and drained,
    // or nullopt with *closed = false on timeout (the `select` of account.go:62-81).
    std::optional<T> recv_until(std::chrono::steady_clock::time_point deadline, bool* closed) {
        std::unique_lock<std::mutex> lk(mu_);
        *closed = false;
        if (!not_empty_.wait_until(lk, deadline, [&] { return !q_.empty() || closed_; })) return std::nullopt;
        if (q_.empty()) { *closed = true; return std::nullopt; }
        T v = std::move(q_.front()); q_.pop_front();
        not_full_.notify_one();
        return v;
    }
    std::optional<T> try_recv() {
        std::lock_guard<std::mutex> lk(mu_);
        if (q_.empty()) return std::nullopt;
        T v = std::move(q_.front()); q_.pop_front();
        not_full_.notify_one();
        return v;
    }
    size_t len() { std::lock_guard<std::mutex> lk(mu_); return q_.size(); }
private:
    std::mutex mu_; std::condition_variable not_empty_, not_full_;
    std::deque<T> q_; size_t cap_; bool closed_ = false;
};

struct Metrics {                                        // the counters account.go:98,120-121 touch
    uint64_t evictions_full = 0, evictions_timeout = 0, evictions_closing = 0, evicted_flows = 0;
    uint64_t buffer_size_gauge = 0;
};

class Accounter {
public:
    using Clock = std::function<uint64_t()>;            // unix ns / monotonic ns
    Accounter(int maxEntries, std::chrono::nanoseconds evictTimeout, Clock clock, Clock monoClock, Metrics* m,
              int device = 0, size_t batchRecords = 1 << 16)
        : evictTimeout_(evictTimeout), clock_(std::move(clock)), monoClock_(std::move(monoClock)), metrics_(m),
          batchCap_(batchRecords) {
        fa_config cfg{};
        cfg.abi_version = FA_ABI_VERSION; cfg.device = device; cfg.mode = FA_MODE_ACCOUNTER;
        cfg.max_entries = (uint64_t)maxEntries; cfg.max_batch = batchRecords;
        if (int rc = fa_create(&cfg, &eng_); rc != 0)
            throw std::runtime_error(std::string("fa_create: ") + fa_last_error());   // no CPU fallback
        batch_.reserve(batchCap_);
    }
    ~Accounter() { fa_destroy(eng_); }
    Accounter(const Accounter&) = delete;

    // Runs until `in` is closed (the goroutine body of account.go:58-100).
    void Account(Chan<RawRecord>& in, Chan<std::vector<Record>>& out) {
        auto nextTick = std::chrono::steady_clock::now() + evictTimeout_;
        for (;;) {
            bool closed = false;
            auto rec = in.recv_until(nextTick, &closed);
            if (rec) {
                batch_.push_back(*rec);
                if (batch_.size() == batchCap_) flush(out, nextTick);
                // Go's select keeps servicing evictTick.C under sustained input (account.go:62-81): do not let a queue
                // that is never empty starve the timeout eviction
                if (std::chrono::steady_clock::now() >= nextTick) {
                    flush(out, nextTick);
                    nextTick = std::chrono::steady_clock::now() + evictTimeout_;
                    size_t live = 0; fa_live_flows(eng_, &live);
                    if (live != 0) evict(out, "timeout", false);
                }
            } else if (closed) {                                  // account.go:73-80
                flush(out, nextTick);
                evict(out, "closing", true);
                return;
            } else {                                              // evictTick.C, account.go:63-71
                flush(out, nextTick);
                nextTick = std::chrono::steady_clock::now() + evictTimeout_;
                size_t live = 0; fa_live_flows(eng_, &live);
                if (live != 0) evict(out, "timeout", false);
            }
            if (metrics_) { size_t live = 0; if (batch_.empty()) { fa_live_flows(eng_, &live); metrics_->buffer_size_gauge = live; } }
        }
    }

private:
    void flush(Chan<std::vector<Record>>& out, std::chrono::steady_clock::time_point& nextTick) {
        size_t off = 0;
        while (off < batch_.size()) {
            size_t took = 0;
            int rc = fa_ingest(eng_, batch_.data() + off, batch_.size() - off, &took);
            if (rc < 0) throw std::runtime_error(std::string("fa_ingest: ") + fa_last_error());
            off += took;
            if (rc == FA_FULL) {                                  // account.go:85-94: evict, reset the ticker
                evict(out, "full", true);
                nextTick = std::chrono::steady_clock::now() + evictTimeout_;
            }
        }
        batch_.clear();
    }
    void evict(Chan<std::vector<Record>>& out, const char* reason, bool evenIfEmpty) {
        const uint64_t now = clock_(), monoNow = monoClock_();   // account.go:103-104
        size_t live = 0; fa_live_flows(eng_, &live);
        std::vector<fa_flow_record> flows(live ? live : 1);
        size_t n = 0;
        if (int rc = fa_evict(eng_, flows.data(), nullptr, nullptr, nullptr, flows.size(), &n); rc != 0)
            throw std::runtime_error(std::string("fa_evict: ") + fa_last_error());
        if (n == 0 && !evenIfEmpty) return;
        std::vector<Record> records; records.reserve(n);
        for (size_t i = 0; i < n; i++) records.push_back(NewRecord(flows[i].id, flows[i].metrics, now, monoNow));
        if (metrics_) {
            if (!strcmp(reason, "full")) metrics_->evictions_full++;
            else if (!strcmp(reason, "timeout")) metrics_->evictions_timeout++;
            else metrics_->evictions_closing++;
            metrics_->evicted_flows += n;
        }
        out.send(std::move(records));                             // evictor <- records (account.go:123)
    }

    fa_engine* eng_ = nullptr;
    std::chrono::nanoseconds evictTimeout_;
    Clock clock_, monoClock_;
    Metrics* metrics_;
    size_t batchCap_;
    std::vector<RawRecord> batch_;
};

inline std::unique_ptr<Accounter> NewAccounter(int maxEntries, std::chrono::nanoseconds evictTimeout, Accounter::Clock clock,
                                               Accounter::Clock monoClock, Metrics* m) {
    return std::make_unique<Accounter>(maxEntries, evictTimeout, std::move(clock), std::move(monoClock), m);
}

}  // namespace flowagg
