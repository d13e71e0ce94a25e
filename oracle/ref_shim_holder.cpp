// oracle/ref_shim_holder.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's scan assembly (ScanDataHolder::pushScanNodeData / rewindCurrentScanData,
// reference src/sdk/src/sl_lidar_driver.cpp:237-371) is a class template private to
// sl_lidar_driver.cpp.  To run the REAL code, this shim compiles that translation unit in place
// (an #include of the file where it lies under /root/reference; nothing is copied) and drives the
// holder exactly as SlamtecLidarDriver does: onHQNodeDecoded -> pushScanNodeData (:1645-1649),
// onHQNodeScanResetReq -> rewindCurrentScanData (:1651-1653), grab -> waitAndLockAvailableScan.
// Built by oracle/Makefile into oracle/_ref/libref_holder.so together with the rest of the SDK
// (minus the separately compiled sl_lidar_driver.cpp).
#include "sl_lidar_driver.cpp"  // -I /root/reference/src/sdk/src

#include <cstring>

extern "C" {

// nodes[0..n): the decoded node stream; resets[0..n_resets): sorted node positions before which
// a scan reset was requested.  Every published scan is appended to scans_out
// (scan k: scan_len[k] nodes at scans_out + k * scan_stride).  Returns the number of scans.
// node_ts (nullable): the timestamp handed over with every node; scan_ts (nullable): receives the
// scan-begin timestamp the holder reports with each published scan (waitAndLockAvailableScan's
// out_timestamp_uS, what grabScanDataHqWithTimeStamp returns).
int ref_assemble_scans_ts(const void* nodes_v, size_t n, const uint32_t* resets, size_t n_resets, size_t max_nodes,
                          void* scans_out_v, size_t scan_stride, uint32_t* scan_len, size_t max_scans,
                          const uint64_t* node_ts, uint64_t* scan_ts) {
  using Node = sl_lidar_response_measurement_node_hq_t;
  const Node* nodes = static_cast<const Node*>(nodes_v);
  Node* scans_out = static_cast<Node*>(scans_out_v);
  sl::ScanDataHolder<Node> holder(max_nodes);
  size_t ri = 0;
  int n_scans = 0;
  for (size_t i = 0; i <= n; ++i) {
    while (ri < n_resets && resets[ri] == i) {
      holder.rewindCurrentScanData();
      ++ri;
    }
    if (i == n) break;
    holder.pushScanNodeData(node_ts ? node_ts[i] : 0, &nodes[i]);
    if (holder.checkNewScanSignalAndReset()) {
      _u64 begin_ts = 0;
      std::vector<Node>* v = holder.waitAndLockAvailableScan(0, &begin_ts);
      if (v) {
        if (static_cast<size_t>(n_scans) < max_scans && v->size() <= scan_stride) {
          std::memcpy(scans_out + static_cast<size_t>(n_scans) * scan_stride, v->data(), v->size() * sizeof(Node));
          scan_len[n_scans] = static_cast<uint32_t>(v->size());
          if (scan_ts) scan_ts[n_scans] = begin_ts;
        }
        ++n_scans;
        holder.unlockScan(v);
      }
    }
  }
  return n_scans;
}

int ref_assemble_scans(const void* nodes_v, size_t n, const uint32_t* resets, size_t n_resets, size_t max_nodes,
                       void* scans_out_v, size_t scan_stride, uint32_t* scan_len, size_t max_scans) {
  return ref_assemble_scans_ts(nodes_v, n, resets, n_resets, max_nodes, scans_out_v, scan_stride, scan_len,
                               max_scans, nullptr, nullptr);
}

}  // extern "C"
