// script_mirror.cpp — TEST INFRASTRUCTURE: a C++17 driver that reads like the reference's spark-shell script, written
// against include/bigclam_b200.hpp (the compiled host-side mirror of the script surface).  Modes:
//
//   script_mirror kset  <minCom> <maxCom> <divCom>                       prints Kset()                  (:116-133, KAT :268)
//   script_mirror steps <edge list> <K> <calls> <F0.f64> <out.bin> <numGPUs>
//        K.set; F = F0; `calls` x backtrackingLineSearchs(uset = all)    trace = LLH of every call      (:152-223)
//   script_mirror sgd   <edge list> <K> <max calls> <F0.f64> <out.bin> <numGPUs>
//        SGDFindC()                                                                                       (:225-243)
//   script_mirror mbsgd <edge list> <K> <max calls> <F0.f64> <out.bin> <numGPUs> <2|3>
//        MBSGD() of bigclamv3-7.scala:206-222 / Bigclamv2.scala:203-219
//   script_mirror sweep <edge list> <minCom> <maxCom> <divCom> <max calls per K>
//        conductanceLocalMin(); for K in Kset: initNeighborComF(K); SGDFindC(); stop rule of :259         (:244-266)
//        prints "K LLH" per K and "KforC k"
//   script_mirror extract <edge list> <K> <calls> <F0.f64> <count>
//        `calls` hot-path calls, then the extraction of Bigclamv2.scala:223-230 with delta from (N, count)
//        prints "delta d" and one line per non-empty community: "c: u u u ..."
//   out.bin: int64 n, k, calls, ntrace | double llh | double trace[ntrace] | double sumF[k] | double F[n*k]
// Exit code 1 with the exception's message on stderr (what the JVM would print as an uncaught exception).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "bigclam_b200.hpp"

static std::vector<double> read_f64(const char *path, size_t count) {
    std::vector<double> v(count);
    std::ifstream fh(path, std::ios::binary);
    if (!fh.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(count * sizeof(double)))) throw bigclam::Error(BIGCLAM_EIO, path);
    return v;
}

static void write_out(const char *path, bigclam::BigClam &b, double llh, const std::vector<double> &trace) {
    const std::vector<double> F = b.F(), sumF = b.sumF();
    const int64_t head[4] = {b.n(), (int64_t)b.K(), b.last_calls, (int64_t)trace.size()};
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char *>(head), sizeof(head));
    out.write(reinterpret_cast<const char *>(&llh), sizeof(double));
    out.write(reinterpret_cast<const char *>(trace.data()), (std::streamsize)(trace.size() * sizeof(double)));
    out.write(reinterpret_cast<const char *>(sumF.data()), (std::streamsize)(sumF.size() * sizeof(double)));
    out.write(reinterpret_cast<const char *>(F.data()), (std::streamsize)(F.size() * sizeof(double)));
    if (!out) throw bigclam::Error(BIGCLAM_EIO, path);
}

int main(int argc, char **argv) try {
    const std::string mode = argc > 1 ? argv[1] : "";
    if (mode == "kset" && argc == 5) {
        for (int k : bigclam::Kset(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]))) std::printf("%d ", k);
        std::printf("\n");
        return 0;
    }
    if (mode == "sweep" && argc == 7) {
        bigclam::BigClam b;
        b.minCom = atoi(argv[3]);
        b.maxCom = atoi(argv[4]);
        b.divCom = atoi(argv[5]);
        b.load_edge_list(argv[2]);
        b.conductanceLocalMin();
        std::vector<std::pair<int, double>> hist;
        const int KforC = b.sweep_K(&hist, 0.001, atoll(argv[6]));
        for (const auto &h : hist) std::printf("%d LLH: %.17g\n", h.first, h.second);          // :258
        std::printf("KforC %d\n", KforC);
        return 0;
    }
    if ((mode == "steps" || mode == "sgd" || mode == "mbsgd") && argc >= 8) {
        const int K = atoi(argv[3]);
        const int64_t calls = atoll(argv[4]);
        bigclam::BigClam b(atoi(argv[7]));
        b.load_edge_list(argv[2]).set_K(K);
        b.set_F(read_f64(argv[5], (size_t)b.n() * (size_t)K));
        double llh = 0.0;
        std::vector<double> trace;
        if (mode == "steps") {
            for (int64_t it = 0; it < calls; ++it) {
                llh = b.backtrackingLineSearchs();
                trace.push_back(llh);
                std::printf("call %lld LLH: %.17g updated %lld\n", (long long)it, llh, (long long)b.last_n_updated);
            }
            b.last_calls = calls;
        } else if (mode == "sgd") {
            llh = b.SGDFindC(1e-4, calls);
            trace = b.last_trace;
            for (size_t i = 1; i < trace.size(); ++i) std::printf("-------Inter: %zu LLH: %.17g\n", i, trace[i]);   // :236
        } else {
            b.MBSGD(argc > 8 ? atoi(argv[8]) : 3, 1e-4, calls);
            trace = b.last_trace;
            llh = trace.empty() ? 0.0 : trace.back();
        }
        write_out(argv[6], b, llh, trace);
        return 0;
    }
    if (mode == "extract" && argc == 7) {
        const int K = atoi(argv[3]);
        bigclam::BigClam b;
        b.load_edge_list(argv[2]).set_K(K);
        b.set_F(read_f64(argv[5], (size_t)b.n() * (size_t)K));
        for (int64_t it = 0; it < atoll(argv[4]); ++it) b.backtrackingLineSearchs();
        const double delta = bigclam::BigClam::delta_threshold(b.n(), atoll(argv[6]));
        std::printf("delta %.17g\n", delta);
        const auto comms = b.extract(delta);
        for (size_t c = 0; c < comms.size(); ++c) {
            if (comms[c].empty()) continue;
            std::printf("%zu:", c);
            for (int32_t u : comms[c]) std::printf(" %d", (int)u);
            std::printf("\n");
        }
        return 0;
    }
    std::fprintf(stderr, "usage: see the header of tests/cpp_host/script_mirror.cpp\n");
    return 2;
} catch (const std::exception &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
}
