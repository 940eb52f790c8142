// host_numa.cpp -- see host_numa.h.  Plain Linux system calls: no libnuma in the image.
#include "host_numa.h"
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

namespace avdsp {

namespace {
constexpr int kMpolPreferred = 1, kMpolBind = 2;        // linux/mempolicy.h
constexpr unsigned kMpolMfMove = 1u << 1;

std::string readFile(const std::string& path) {
    std::ifstream f(path);
    if (!f) return std::string();
    std::stringstream ss; ss << f.rdbuf();
    return ss.str();
}
} // namespace

std::vector<int> parseCpuList(const std::string& text) {
    std::vector<int> cpus;
    size_t i = 0;
    while (i < text.size()) {
        while (i < text.size() && !isdigit((unsigned char)text[i])) i++;
        if (i >= text.size()) break;
        int a = 0;
        while (i < text.size() && isdigit((unsigned char)text[i])) a = a * 10 + (text[i++] - '0');
        int b = a;
        if (i < text.size() && text[i] == '-') {
            i++; b = 0;
            while (i < text.size() && isdigit((unsigned char)text[i])) b = b * 10 + (text[i++] - '0');
        }
        for (int c = a; c <= b && c < 4096; c++) cpus.push_back(c);
    }
    return cpus;
}

int numaNodeOfPci(const char* busId) {
    if (!busId || !*busId) return -1;
    std::string id(busId);
    for (char& c : id) c = (char)tolower((unsigned char)c);
    // CUDA prints an 8-digit domain ("00000000:1B:00.0"); sysfs uses four
    const size_t colon = id.find(':');
    if (colon != std::string::npos && colon > 4) id = id.substr(colon - 4);
    const std::string t = readFile("/sys/bus/pci/devices/" + id + "/numa_node");
    if (t.empty()) return -1;
    const int n = atoi(t.c_str());
    return n >= 0 ? n : -1;
}

std::vector<int> cpusOfNode(int node) {
    if (node < 0) return {};
    return parseCpuList(readFile("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist"));
}

bool bindThreadToNode(int node) {
    const std::vector<int> cpus = cpusOfNode(node);
    if (cpus.empty()) return false;
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed); CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return false;
    int n = 0;
    for (int c : cpus) if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); n++; }
    if (n == 0) return false;                         // the container's cpuset has none of that node's CPUs
    if (sched_setaffinity(0, sizeof want, &want) != 0) return false;
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof mask * 8)) {
        mask[node / (8 * sizeof(long))] |= 1ul << (node % (8 * sizeof(long)));
        syscall(SYS_set_mempolicy, kMpolPreferred, mask, sizeof mask * 8);     // best effort
    }
    return true;
}

void* hostAllocPlaced(size_t bytes, const std::vector<size_t>& off, const std::vector<int>& nodes) {
    if (bytes == 0) bytes = 1;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    char* base = (char*)p;
    for (size_t k = 0; k + 1 < off.size() && k < nodes.size(); k++) {
        const int node = nodes[k];
        if (node < 0) continue;
        // whole pages inside the range; a page shared by two ranges keeps the default policy
        const size_t a = (off[k] + page - 1) / page * page, b = std::min(off[k + 1], len) / page * page;
        if (b <= a) continue;
        unsigned long mask[16] = {0};
        if (node >= (int)(sizeof mask * 8)) continue;
        mask[node / (8 * sizeof(long))] |= 1ul << (node % (8 * sizeof(long)));
        syscall(SYS_mbind, base + a, b - a, kMpolBind, mask, sizeof mask * 8, kMpolMfMove);             // best effort
    }
    // first touch (pages get their node here); one byte per page
    for (size_t o = 0; o < len; o += page) base[o] = 0;
    return p;
}

void hostFreePlaced(void* p, size_t bytes) {
    if (!p) return;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    munmap(p, (std::max<size_t>(bytes, 1) + page - 1) / page * page);
}

} // namespace avdsp
