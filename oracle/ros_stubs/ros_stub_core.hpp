// oracle/ros_stubs/ros_stub_core.hpp -- TEST INFRASTRUCTURE ONLY.
//
// ROS 2 is not installed in this image, so the reference's node (src/rplidar_node.cpp) cannot be built
// against the real rclcpp.  These headers declare JUST ENOUGH of the ROS 2 API -- types, signatures and
// do-nothing bodies, written from the public API documentation, no ROS source involved -- for that file to
// compile unmodified, in place, so that the REAL RPlidarNode::publish_scan (src/rplidar_node.cpp:556-680)
// can be executed as the oracle of the LaserScan arithmetic (oracle/ref_shim_node.cpp).  Nothing here
// implements ROS behaviour: publishers hand the message to a capture hook, parameters return their
// declared defaults, logging is dropped, time is a plain nanosecond counter.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <variant>
#include <vector>

// ---- messages ------------------------------------------------------------------------------------------------
namespace rclcpp {
class Time;
}
namespace builtin_interfaces::msg {
struct Time {
  int32_t sec = 0;
  uint32_t nanosec = 0;
  Time() = default;
  Time(const rclcpp::Time& t);  // NOLINT: header.stamp = node->now()
};
}  // namespace builtin_interfaces::msg
namespace std_msgs::msg {
struct Header {
  builtin_interfaces::msg::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs::msg
namespace sensor_msgs::msg {
struct LaserScan {
  std_msgs::msg::Header header;
  float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0, range_min = 0,
        range_max = 0;
  std::vector<float> ranges, intensities;
};
struct PointField {
  static constexpr uint8_t INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8;
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 0;
};
struct PointCloud2 {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
}  // namespace sensor_msgs::msg
namespace geometry_msgs::msg {
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
};
struct Transform {
  Vector3 translation;
  Quaternion rotation;
};
struct TransformStamped {
  std_msgs::msg::Header header;
  std::string child_frame_id;
  Transform transform;
};
}  // namespace geometry_msgs::msg
namespace rcl_interfaces::msg {
struct SetParametersResult {
  bool successful = true;
  std::string reason;
};
}  // namespace rcl_interfaces::msg
namespace lifecycle_msgs::msg {
struct State {
  static constexpr uint8_t PRIMARY_STATE_UNKNOWN = 0, PRIMARY_STATE_UNCONFIGURED = 1, PRIMARY_STATE_INACTIVE = 2,
                           PRIMARY_STATE_ACTIVE = 3, PRIMARY_STATE_FINALIZED = 4;
};
}  // namespace lifecycle_msgs::msg
namespace diagnostic_msgs::msg {
struct DiagnosticStatus {
  static constexpr uint8_t OK = 0, WARN = 1, ERROR = 2, STALE = 3;
};
}  // namespace diagnostic_msgs::msg

// ---- rclcpp ----------------------------------------------------------------------------------------------------
namespace rclcpp {
class NodeOptions {};
class Duration {
 public:
  explicit Duration(int64_t ns = 0) : ns_(ns) {}
  double seconds() const { return static_cast<double>(ns_) * 1e-9; }
  int64_t nanoseconds() const { return ns_; }

 private:
  int64_t ns_;
};
class Time {
 public:
  Time() = default;
  explicit Time(int64_t ns) : ns_(ns) {}
  int64_t nanoseconds() const { return ns_; }
  double seconds() const { return static_cast<double>(ns_) * 1e-9; }
  Duration operator-(const Time& o) const { return Duration(ns_ - o.ns_); }

 private:
  int64_t ns_ = 0;
};
class Clock {
 public:
  using SharedPtr = std::shared_ptr<Clock>;
};
class Logger {};
enum class ParameterType { PARAMETER_NOT_SET, PARAMETER_BOOL, PARAMETER_INTEGER, PARAMETER_DOUBLE, PARAMETER_STRING };
class Parameter {
 public:
  using Value = std::variant<bool, int64_t, double, std::string>;
  Parameter() = default;
  Parameter(std::string name, Value v) : name_(std::move(name)), v_(std::move(v)) {}
  const std::string& get_name() const { return name_; }
  ParameterType get_type() const {
    switch (v_.index()) {
      case 0: return ParameterType::PARAMETER_BOOL;
      case 1: return ParameterType::PARAMETER_INTEGER;
      case 2: return ParameterType::PARAMETER_DOUBLE;
      default: return ParameterType::PARAMETER_STRING;
    }
  }
  bool as_bool() const { return std::get<bool>(v_); }
  int64_t as_int() const { return std::get<int64_t>(v_); }
  double as_double() const { return std::get<double>(v_); }
  const std::string& as_string() const { return std::get<std::string>(v_); }

 private:
  std::string name_;
  Value v_;
};
class QoS {
 public:
  explicit QoS(size_t) {}
  QoS& reliable() { return *this; }
  QoS& best_effort() { return *this; }
  QoS& durability_volatile() { return *this; }
  QoS& transient_local() { return *this; }
  QoS& keep_last(size_t) { return *this; }
};
class SensorDataQoS : public QoS {
 public:
  SensorDataQoS() : QoS(5) {}
};
inline bool ok() { return true; }
inline void init(int, char**) {}
inline void shutdown() {}
namespace node_interfaces {
class NodeBaseInterface {
 public:
  using SharedPtr = std::shared_ptr<NodeBaseInterface>;
};
}  // namespace node_interfaces
namespace executors {
class MultiThreadedExecutor {
 public:
  template <class T>
  void add_node(T&&) {}
  void spin() {}
};
class SingleThreadedExecutor : public MultiThreadedExecutor {};
}  // namespace executors
}  // namespace rclcpp
inline builtin_interfaces::msg::Time::Time(const rclcpp::Time& t)
    : sec(static_cast<int32_t>(t.nanoseconds() / 1000000000)),
      nanosec(static_cast<uint32_t>(t.nanoseconds() % 1000000000)) {}

// logging is dropped; the arguments are still evaluated for their types
template <class... A>
inline void ros_stub_log(A&&...) {}
#define RCLCPP_DEBUG(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_INFO(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_WARN(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_ERROR(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_FATAL(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_INFO_ONCE(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_WARN_ONCE(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_DEBUG_THROTTLE(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_INFO_THROTTLE(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_WARN_THROTTLE(...) ros_stub_log(__VA_ARGS__)
#define RCLCPP_ERROR_THROTTLE(...) ros_stub_log(__VA_ARGS__)

// ---- rclcpp_lifecycle ------------------------------------------------------------------------------------------
namespace ros_stub {
// where LifecyclePublisher<LaserScan>::publish leaves the message (set by oracle/ref_shim_node.cpp)
inline std::function<void(const sensor_msgs::msg::LaserScan&)>& laserscan_sink() {
  static thread_local std::function<void(const sensor_msgs::msg::LaserScan&)> f;  // one per worker thread
  return f;
}
inline std::function<void(const sensor_msgs::msg::PointCloud2&)>& pointcloud2_sink() {
  static thread_local std::function<void(const sensor_msgs::msg::PointCloud2&)> f;
  return f;
}
template <class Msg>
inline void deliver(const Msg&) {}
inline void deliver(const sensor_msgs::msg::LaserScan& m) {
  if (laserscan_sink()) laserscan_sink()(m);
}
inline void deliver(const sensor_msgs::msg::PointCloud2& m) {
  if (pointcloud2_sink()) pointcloud2_sink()(m);
}
}  // namespace ros_stub

namespace rclcpp_lifecycle {
class State {
 public:
  explicit State(uint8_t id = lifecycle_msgs::msg::State::PRIMARY_STATE_UNCONFIGURED) : id_(id) {}
  uint8_t id() const { return id_; }
  std::string label() const { return "stub"; }

 private:
  uint8_t id_;
};
namespace node_interfaces {
class LifecycleNodeInterface {
 public:
  enum class CallbackReturn : uint8_t { SUCCESS = 0, FAILURE = 1, ERROR = 2 };
  virtual ~LifecycleNodeInterface() = default;
  virtual CallbackReturn on_configure(const State&) { return CallbackReturn::SUCCESS; }
  virtual CallbackReturn on_cleanup(const State&) { return CallbackReturn::SUCCESS; }
  virtual CallbackReturn on_shutdown(const State&) { return CallbackReturn::SUCCESS; }
  virtual CallbackReturn on_activate(const State&) { return CallbackReturn::SUCCESS; }
  virtual CallbackReturn on_deactivate(const State&) { return CallbackReturn::SUCCESS; }
  virtual CallbackReturn on_error(const State&) { return CallbackReturn::SUCCESS; }
};
}  // namespace node_interfaces

template <class Msg>
class LifecyclePublisher {
 public:
  using SharedPtr = std::shared_ptr<LifecyclePublisher<Msg>>;
  void publish(const Msg& m) { ros_stub::deliver(m); }
  void publish(std::unique_ptr<Msg> m) {
    if (m) ros_stub::deliver(*m);
  }
  void on_activate() { active_ = true; }
  void on_deactivate() { active_ = false; }
  bool is_activated() const { return active_; }

 private:
  bool active_ = true;
};

class LifecycleNode : public node_interfaces::LifecycleNodeInterface {
 public:
  struct OnSetParametersCallbackHandle {
    using SharedPtr = std::shared_ptr<OnSetParametersCallbackHandle>;
    std::function<rcl_interfaces::msg::SetParametersResult(const std::vector<rclcpp::Parameter>&)> callback;
  };
  LifecycleNode(const std::string& name, const rclcpp::NodeOptions& = rclcpp::NodeOptions()) : name_(name) {}

  template <class T>
  T declare_parameter(const std::string& name, const T& def) {
    if (overridden_.count(name)) {  // an override given before the declaration wins, as in ROS
      T v{};
      get_parameter(name, v);
      return v;
    }
    set_default(name, def);
    return def;
  }
  bool has_parameter(const std::string& name) const { return params_.count(name) != 0; }
  // test hook (not ROS API): what a launch file / `ros2 param set` would do before the node reads the value
  template <class T>
  void stub_override_parameter(const std::string& name, const T& v) { set_default(name, v); overridden_[name] = true; }
  template <class T>
  bool get_parameter(const std::string& name, T& out) const {
    auto it = params_.find(name);
    if (it == params_.end()) return false;
    convert(it->second, out);
    return true;
  }
  template <class T>
  bool get_parameter_or(const std::string& name, T& out, const T& alt) const {
    if (get_parameter(name, out)) return true;
    out = alt;
    return false;
  }
  template <class F>
  OnSetParametersCallbackHandle::SharedPtr add_on_set_parameters_callback(F&& f) {
    auto h = std::make_shared<OnSetParametersCallbackHandle>();
    h->callback = std::forward<F>(f);
    return h;
  }
  template <class Msg>
  typename LifecyclePublisher<Msg>::SharedPtr create_publisher(const std::string&, const rclcpp::QoS&) {
    return std::make_shared<LifecyclePublisher<Msg>>();
  }
  rclcpp::Time now() const { return rclcpp::Time(0); }
  rclcpp::Logger get_logger() const { return rclcpp::Logger(); }
  rclcpp::Clock::SharedPtr get_clock() const { return std::make_shared<rclcpp::Clock>(); }
  State get_current_state() const { return State(lifecycle_msgs::msg::State::PRIMARY_STATE_ACTIVE); }
  rclcpp::node_interfaces::NodeBaseInterface::SharedPtr get_node_base_interface() const {
    return std::make_shared<rclcpp::node_interfaces::NodeBaseInterface>();
  }
  const char* get_name() const { return name_.c_str(); }

 private:
  using Stored = std::variant<bool, int64_t, double, std::string>;
  void set_default(const std::string& n, bool v) { params_[n] = v; }
  void set_default(const std::string& n, int v) { params_[n] = static_cast<int64_t>(v); }
  void set_default(const std::string& n, int64_t v) { params_[n] = v; }
  void set_default(const std::string& n, float v) { params_[n] = static_cast<double>(v); }
  void set_default(const std::string& n, double v) { params_[n] = v; }
  void set_default(const std::string& n, const std::string& v) { params_[n] = v; }
  void set_default(const std::string& n, const char* v) { params_[n] = std::string(v); }
  static void convert(const Stored& s, bool& o) { o = std::get<bool>(s); }
  static void convert(const Stored& s, int& o) { o = static_cast<int>(std::get<int64_t>(s)); }
  static void convert(const Stored& s, int64_t& o) { o = std::get<int64_t>(s); }
  static void convert(const Stored& s, float& o) { o = static_cast<float>(std::get<double>(s)); }
  static void convert(const Stored& s, double& o) { o = std::get<double>(s); }
  static void convert(const Stored& s, std::string& o) { o = std::get<std::string>(s); }
  std::string name_;
  std::map<std::string, Stored> params_;
  std::map<std::string, bool> overridden_;
};
}  // namespace rclcpp_lifecycle

// ---- tf2 / tf2_ros / diagnostic_updater --------------------------------------------------------------------------
namespace tf2 {
class Quaternion {
 public:
  void setRPY(double, double, double) {}
  double x() const { return 0; }
  double y() const { return 0; }
  double z() const { return 0; }
  double w() const { return 1; }
};
}  // namespace tf2
namespace tf2_ros {
class StaticTransformBroadcaster {
 public:
  template <class NodeT>
  explicit StaticTransformBroadcaster(NodeT&&) {}
  void sendTransform(const geometry_msgs::msg::TransformStamped&) {}
};
}  // namespace tf2_ros
namespace diagnostic_updater {
class DiagnosticStatusWrapper {
 public:
  void summary(uint8_t, const std::string&) {}
  template <class T>
  void add(const std::string&, const T&) {}
  template <class... A>
  void addf(const std::string&, const char*, A&&...) {}
};
class Updater {
 public:
  template <class NodeT>
  explicit Updater(NodeT&&) {}
  void setHardwareID(const std::string&) {}
  template <class T>
  void add(const std::string&, T*, void (T::*)(DiagnosticStatusWrapper&)) {}
  void force_update() {}
  void update() {}
};
}  // namespace diagnostic_updater
