/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU oracle (liboracle_scan.so).  Only tests/, the
 * __graft_entry__.smoke() check and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (librplidar_b200.so) never links or calls it.
 *
 * Part 1 (scan_oracle.cpp) restates the reference's per-scan hot path:
 *   - ascendScanData_<hq>              reference src/sdk/src/sl_lidar_driver.cpp:128-184
 *   - RPlidarNode::publish_scan body   reference src/rplidar_node.cpp:556-680
 * PARITY PINNED: orc_ascend_scan is checked node-for-node against the reference's own
 * compiled ascendScanData (oracle/_ref, built by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, and both are checked against golden vectors captured
 * from the reference's DummyLidarDriver (tests/golden/, made by
 * tests/golden/make_golden.py).  publish_scan itself needs rclcpp (absent here), so its
 * restatement is pinned by line-by-line citation plus the golden LaserScan arrays.
 *
 * Part 2 (cloud_oracle.cpp) defines the north-star extensions that have NO reference
 * implementation (polar->xyz, PointCloud2 packing, range/intensity window, voxel grid,
 * statistical outlier removal).  PARITY UNPINNED: self-authored definitions; see the
 * header of cloud_oracle.cpp.
 */
#ifndef RPL_ORACLE_H_
#define RPL_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/sdk/include/sl_lidar_cmd.h:272-278 : 8 bytes, dist at unaligned offset 2 */
typedef struct __attribute__((packed)) orc_node_hq {
  uint16_t angle_z_q14;
  uint32_t dist_mm_q2;
  uint8_t quality;
  uint8_t flag;
} orc_node_hq;

#define ORC_RESULT_OK 0u
#define ORC_RESULT_OPERATION_FAIL 0x80008001u /* reference sl_types.h: SL_RESULT_OPERATION_FAIL */

/* LaserScan scalar fields filled by publish_scan (reference rplidar_node.cpp:616-625,
 * :632-636, :664-668). */
typedef struct orc_scan_header {
  float angle_min;
  float angle_max;
  float angle_increment;
  float time_increment;
  float scan_time;
  float range_min;
  float range_max;
  uint32_t beam_count; /* ranges.size() */
  uint32_t published;  /* 0 when the reference returns early (no valid point) */
} orc_scan_header;

typedef struct orc_scan_params {
  uint8_t is_new_protocol; /* reference rplidar_node.cpp:575-579 */
  uint8_t scan_processing; /* Mode A (1) / Mode B (0), :630 */
  uint8_t inverted;        /* :644, :673 */
  uint8_t apply_ascend;    /* angle_compensate -> profile_.apply_geometric_correction,
                              reference lidar_driver_wrapper.cpp:107,:328 */
  float range_max;         /* cached_current_max_range_, :389-394 */
  double scan_duration;    /* seconds, :442 */
} orc_scan_params;

/* ---- part 1: the reference path ------------------------------------------------ */

/* In place.  stable=0: std::sort exactly as the reference (tie order = libstdc++
 * introsort).  stable=1: std::stable_sort -- the documented tie rule of the CUDA path. */
uint32_t orc_ascend_scan(orc_node_hq* nodes, size_t count, int stable);

/* ranges/intensities must hold `count` floats.  Returns header.published. */
uint32_t orc_publish_scan(const orc_node_hq* nodes, size_t count, const orc_scan_params* p,
                          int stable, float* ranges, float* intensities, orc_scan_header* hdr);

/* grab_scan_data glue + publish_scan for a batch (one scan per task, `threads` workers):
 * nodes are modified in place when p->apply_ascend (like the wrapper's buffer).
 * Outputs are strided like the inputs.  Returns wall seconds of the timed region. */
double orc_pipeline_batch(orc_node_hq* nodes, const uint32_t* counts, uint32_t n_scans,
                          uint32_t stride, const orc_scan_params* p, int stable, float* ranges,
                          float* intensities, uint32_t* beam_counts, float* angle_increment,
                          uint32_t* status, int threads);

/* Restatement of DummyLidarDriver::grab_scan_data's generator for call number
 * `call_index` (1-based: the reference pre-increments phase by 0.1f per call).
 * reference src/lidar_driver_wrapper.cpp:441-471.  Writes 360 nodes. */
void orc_dummy_scan(uint32_t call_index, orc_node_hq* out360);

/* Deterministic synthetic scans of SURVEY.md 8(d) (splitmix64, seed 0x5EED0000+scan_id).
 * variant: 0 = C2 tie-free rotated (5% invalid, quality 188), 1 = same with U[0,255]
 * quality, 2 = tie variant (iid U[0,65535] keys), 3 = C4 (tie-free keys, iid shuffled),
 * 4 = C3 "room" (variant 0 keys, 16 constant-range arcs of 2..10 m + 2 cm noise). */
void orc_synth_scan(uint64_t scan_id, uint32_t n, int variant, orc_node_hq* out);
void orc_synth_batch(uint64_t first_scan_id, uint32_t n_scans, uint32_t n, uint32_t stride,
                     int variant, orc_node_hq* out, int threads);

/* ---- part 1b: dense-capsule decode (decode_oracle.cpp) ------ */
#define ORC_DENSE_CAPSULE_BYTES 84 /* reference sl_lidar_cmd.h:228-234 */
#define ORC_CAPSULE_OK 1u                /* sync nibbles and checksum fine */
#define ORC_CAPSULE_SYNC 2u              /* start_angle_sync_q6 bit 15: first capsule of a revolution */
#define ORC_CAPSULE_EMIT 4u              /* this capsule released the previous capsule's 40 nodes */
#define ORC_CAPSULE_DISCARD 8u           /* angular jump above the 100 Hz bound: nothing released */
#define ORC_CAPSULE_CHECKSUM_ERR 16u     /* ERR_EVENT_ON_EXP_CHECKSUM_ERR */
#define ORC_CAPSULE_ENCODER_RESET_ERR 32u /* ERR_EVENT_ON_EXP_ENCODER_RESET (with SYNC) */
#define ORC_CAPSULE_BAD_FRAME 64u        /* wrong sync nibbles: outside the framed contract */
/* Decodes n_capsules framed capsules of one stream.  *sync_state: in/out, the reference's
 * function-static lastNodeSyncBit.  nodes_out must hold 40 * n_capsules nodes.  Returns the
 * number of nodes written; capsule_node_offset[j] = nodes written before capsule j. */
uint32_t orc_dense_decode(const uint8_t* capsules, uint32_t n_capsules, uint32_t sample_duration_us,
                          uint32_t* sync_state, orc_node_hq* nodes_out, uint32_t* capsule_status,
                          uint32_t* capsule_node_offset);

/* Scan assembly: cuts a decoded node stream into scans exactly as
 * ScanDataHolder does.  resets: sorted node positions before which a scan reset was requested.
 * Published scan k goes to scans_out + k * scan_stride (scan_len[k] nodes); at most max_scans are
 * stored.  Returns the number of published scans. */
uint32_t orc_assemble_scans(const orc_node_hq* nodes, uint32_t n, const uint32_t* resets, uint32_t n_resets,
                            uint32_t max_nodes, orc_node_hq* scans_out, uint32_t scan_stride,
                            uint32_t* scan_len, uint32_t max_scans);

/* ---- part 1c: the other measurement answer formats (capsule_oracle.cpp) ---------- */
#define ORC_ANS_NORMAL 0x81u      /* 5-byte standard nodes (raw byte stream) */
#define ORC_ANS_EXPRESS 0x82u     /* 84 B -> 32 nodes */
#define ORC_ANS_HQ 0x83u          /* 781 B -> 96 nodes, CRC32 */
#define ORC_ANS_ULTRA 0x84u       /* 132 B -> 96 nodes */
#define ORC_ANS_DENSE 0x85u       /* 84 B -> 40 nodes (decode_oracle.cpp) */
#define ORC_ANS_ULTRA_DENSE 0x86u /* 170 B -> 64 nodes */
uint32_t orc_capsule_bytes(uint32_t ans_type);
uint32_t orc_capsule_nodes(uint32_t ans_type);
uint32_t orc_crc32_padded(const uint8_t* p, uint32_t len);
/* Framed capsules of any capsule format.  state[2]: in/out {last node sync bit, last distance}
 * (dense / ultra-dense only).  nodes_out must hold orc_capsule_nodes * n_capsules nodes. */
uint32_t orc_frame_capsules(uint32_t ans_type, const uint8_t* bytes, uint32_t n, uint8_t* capsules_out,
                            uint32_t max_capsules, uint32_t* bytes_left);
uint32_t orc_decode_capsules(uint32_t ans_type, const uint8_t* capsules, uint32_t n_capsules,
                             uint32_t sample_duration_us, uint32_t* state, orc_node_hq* nodes_out,
                             uint32_t* capsule_status, uint32_t* capsule_node_offset);
/* Standard nodes from a raw byte stream, with the reference's byte-level resynchronisation. */
uint32_t orc_decode_normal(const uint8_t* bytes, uint32_t n_bytes, orc_node_hq* nodes_out, uint32_t* node_end,
                           uint32_t* fsm_pos);

/* ---- part 1d: per-sample timestamps (timestamp_oracle.cpp) ---- */
/* timing4 = {sample_duration_uS, native_baudrate, linkage_delay_uS, native_interface_type}
 * (sl::SlamtecLidarTimingDesc, reference src/sdk/include/sl_lidar_driver.h:156-166). */
uint64_t orc_sample_delay_us(uint32_t ans_type, const uint32_t* timing4, uint32_t sample_idx);
void orc_node_timestamps(uint32_t ans_type, const uint32_t* timing4, const uint64_t* capsule_rx_us,
                         const uint32_t* capsule_status, const uint32_t* capsule_node_offset, uint32_t n_capsules,
                         uint64_t* ts_out);
void orc_normal_timestamps(const uint32_t* timing4, const uint32_t* node_end, uint32_t n_nodes, uint32_t chunk_bytes,
                           const uint64_t* chunk_rx_us, uint64_t* ts_out);
uint32_t orc_assemble_scans_ts(const orc_node_hq* nodes, uint32_t n, const uint32_t* resets, uint32_t n_resets,
                               uint32_t max_nodes, orc_node_hq* scans_out, uint32_t scan_stride, uint32_t* scan_len,
                               uint32_t max_scans, const uint64_t* node_ts, uint64_t* scan_ts);

/* ---- part 2: extensions (parity unpinned) --------------------------------------- */

typedef struct orc_cloud_params {
  float range_min;      /* keep range_min <= r <= range_max */
  float range_max;
  float intensity_min;  /* keep intensity >= intensity_min */
  float voxel_size;     /* metres; 0 = no voxel grid */
  uint32_t sor_k;       /* 0 = no statistical outlier removal */
  float sor_alpha;
  uint8_t is_new_protocol;
  uint8_t pad[3];
} orc_cloud_params;

/* One scan -> filtered, angle-sorted polar points -> xyz (+intensity), optionally SOR,
 * optionally voxel-grid centroids.  xyzi holds 4 floats per output point (PointCloud2
 * point_step 16: x,y,z,intensity).  Returns the number of output points. */
uint32_t orc_cloud_scan(const orc_node_hq* nodes, size_t count, const orc_cloud_params* p,
                        float* xyzi);

#ifdef __cplusplus
}
#endif
#endif /* RPL_ORACLE_H_ */
