// Section timers with the reference's macro surface and report format (GlobalBenchmark.hh:8-58,
// Timer.hh:39-204): sections nest through a stack, names are colon-joined paths, the report
// prints "<4*depth spaces><leaf>\t<seconds>\t<invocations>" sorted by path, then "Full time".
// Always compiled in (the reference needs -DMESHFEM_ENABLE_BENCHMARKING); device work reports
// CUDA-event time through addDeviceTime().
#ifndef MESHFEM_B200_GLOBALBENCHMARK_HH
#define MESHFEM_B200_GLOBALBENCHMARK_HH
#include <chrono>
#include <iostream>
#include <list>
#include <map>
#include <string>

class Timer {
public:
    struct Entry { double elapsed = 0; size_t invocations = 0; std::chrono::steady_clock::time_point t0; bool running = false; };
    Timer() { reset(); }
    void reset() { m_entries.clear(); m_stack.clear(); m_start = std::chrono::steady_clock::now(); }
    void startSection(const std::string &name) {
        const std::string full = path(name);
        start(full);
        m_stack.push_back(name);
    }
    void stopSection(const std::string &name) {
        if (m_stack.empty() || m_stack.back() != name) { std::cerr << "Timer: section mismatch " << name << std::endl; return; }
        m_stack.pop_back();
        stop(path(name));
    }
    void start(const std::string &name) { startFull(m_stack.empty() || isFull(name) ? name : name); }
    void startTimer(const std::string &name) { startFull(path(name)); }
    void stopTimer(const std::string &name) { stop(path(name)); }
    void addSeconds(const std::string &name, double s) { auto &e = m_entries[path(name)]; e.elapsed += s; e.invocations++; }
    void report(std::ostream &os) const {
        for (const auto &kv : m_entries) os << displayName(kv.first) << '\t' << kv.second.elapsed << '\t' << kv.second.invocations << std::endl;
        os << "Full time\t" << std::chrono::duration<double>(std::chrono::steady_clock::now() - m_start).count() << std::endl;
    }
    const std::map<std::string, Entry> &entries() const { return m_entries; }

private:
    std::map<std::string, Entry> m_entries;
    std::list<std::string> m_stack;
    std::chrono::steady_clock::time_point m_start;
    std::string path(const std::string &leaf) const {
        std::string p;
        for (const auto &s : m_stack) p += s + ":";
        return p + leaf;
    }
    static bool isFull(const std::string &n) { return n.find(':') != std::string::npos; }
    void startFull(const std::string &full) { auto &e = m_entries[full]; e.t0 = std::chrono::steady_clock::now(); e.running = true; e.invocations++; }
    void stop(const std::string &full) {
        auto it = m_entries.find(full);
        if (it == m_entries.end() || !it->second.running) return;
        it->second.elapsed += std::chrono::duration<double>(std::chrono::steady_clock::now() - it->second.t0).count();
        it->second.running = false;
    }
    static std::string displayName(const std::string &name) {
        size_t levels = 0;
        for (char c : name) if (c == ':') ++levels;
        if (levels == 0) return name;
        std::string result(4 * levels, ' ');
        result.append(name, name.rfind(':') + 1, std::string::npos);
        return result;
    }
};

inline Timer &globalTimer() { static Timer t; return t; }

inline void BENCHMARK_START_TIMER_SECTION(const std::string &n) { globalTimer().startSection(n); }
inline void BENCHMARK_STOP_TIMER_SECTION(const std::string &n) { globalTimer().stopSection(n); }
inline void BENCHMARK_START_TIMER(const std::string &n) { globalTimer().startTimer(n); }
inline void BENCHMARK_STOP_TIMER(const std::string &n) { globalTimer().stopTimer(n); }
inline void BENCHMARK_ADD_DEVICE_SECONDS(const std::string &n, double s) { globalTimer().addSeconds(n, s); }
inline void BENCHMARK_REPORT() { globalTimer().report(std::cout); }
inline void BENCHMARK_RESET() { globalTimer().reset(); }
struct BENCHMARK_SCOPED_TIMER_SECTION {
    std::string name;
    explicit BENCHMARK_SCOPED_TIMER_SECTION(const std::string &n) : name(n) { BENCHMARK_START_TIMER_SECTION(n); }
    ~BENCHMARK_SCOPED_TIMER_SECTION() { BENCHMARK_STOP_TIMER_SECTION(name); }
};
#endif
