// host_numa.h -- where a GPU sits in the host's memory topology, and host buffers placed next to it.
//
// The end-to-end rate of the host path is set by the PCIe DMA of the PCM in and out (C2: 7.9 GB per 8.6 ms of kernel
// time).  On a two-socket box a pinned buffer that lives on the other socket's memory crosses the socket interconnect on
// every DMA; eight ranks whose pinned buffers all landed on one node measured 105 GB/s in total.  So the
// staging threads run on the CPUs of their GPU's node and the PCM buffers the library hands out are bound to that node.
// Linux only (sysfs + the mempolicy system calls); every function degrades to "no placement" when the topology is unknown.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace avdsp {

// NUMA node of a PCI device given its "domain:bus:device.function" id (as cudaDeviceGetPCIBusId prints it); -1 when unknown
int numaNodeOfPci(const char* busId);
// CPUs of a node (parsed from /sys/devices/system/node/nodeN/cpulist); empty when unknown
std::vector<int> cpusOfNode(int node);
// parse a sysfs cpulist ("0-15,32-47")
std::vector<int> parseCpuList(const std::string& text);
// run the calling thread on the node's CPUs and prefer its memory for the thread's allocations; false: left as it was
bool bindThreadToNode(int node);
// page-aligned anonymous memory of `bytes`; [off[k], off[k+1]) is bound to nodes[k] (-1: default policy) and touched.
// Returns nullptr on failure.  Free with hostFreePlaced.
void* hostAllocPlaced(size_t bytes, const std::vector<size_t>& off, const std::vector<int>& nodes);
void hostFreePlaced(void* p, size_t bytes);

} // namespace avdsp
