// rclcpp stand-in for oracle/_ref (TEST INFRASTRUCTURE ONLY): just enough of the rclcpp surface for the reference's
// node classes (/root/reference/src/aos_seed_gen_node.cpp:65-227, src/aos_gvd_node.cpp:22-81) to be constructed and
// their callbacks / private members to be called in-process.  Publishers keep the last message per topic in a
// process-wide board the harness reads back; subscriptions and timers are inert (the harness calls the callbacks).
#pragma once
#include <any>
#include <optional>
#include <chrono>
#include <cstdio>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <typeindex>
#include <vector>

#include "ref_shim_msgs.hpp"

namespace ref_shim {
struct Board {  // last message published per topic
  std::map<std::string, std::shared_ptr<void>> last;
  std::map<std::string, int> count;
  static Board &get() { static Board b; return b; }
};
struct ParamOverrides {  // values the harness wants declare_parameter / get_parameter to return
  std::map<std::string, double> num;
  std::map<std::string, std::string> str;
  static ParamOverrides &get() { static ParamOverrides p; return p; }
};
}  // namespace ref_shim

namespace rclcpp {

struct Duration {
  double s = 0;
  static Duration from_seconds(double v) { Duration d; d.s = v; return d; }
  double seconds() const { return s; }
  operator builtin_interfaces::msg::Duration() const {
    builtin_interfaces::msg::Duration d;
    d.sec = (int32_t)s;
    d.nanosec = (uint32_t)((s - (double)d.sec) * 1e9);
    return d;
  }
};
struct Time {
  double t = 0;
  Time() = default;
  explicit Time(double v) : t(v) {}
  Time(const builtin_interfaces::msg::Time &m) : t(m.sec + 1e-9 * m.nanosec) {}
  double seconds() const { return t; }
  Duration operator-(const Time &o) const { return Duration::from_seconds(t - o.t); }
  operator builtin_interfaces::msg::Time() const {
    builtin_interfaces::msg::Time m;
    m.sec = (int32_t)t;
    m.nanosec = (uint32_t)((t - (double)m.sec) * 1e9);
    return m;
  }
};
struct Clock {
  using SharedPtr = std::shared_ptr<Clock>;
  Time now() const { static double t = 1000.0; t += 0.001; return Time(t); }
};
struct Logger {};

enum class DurabilityPolicy { TransientLocal, Volatile, SystemDefault };
enum class HistoryPolicy { KeepLast, KeepAll, SystemDefault };
enum class ReliabilityPolicy { Reliable, BestEffort, SystemDefault };
struct QoS {
  QoS(size_t = 10) {}
  QoS &reliable() { return *this; }
  QoS &best_effort() { return *this; }
  QoS &durability(DurabilityPolicy) { return *this; }
  QoS &history(HistoryPolicy) { return *this; }
  QoS &keep_last(size_t) { return *this; }
  QoS &transient_local() { return *this; }
  QoS &durability_volatile() { return *this; }
};

template <class M>
struct Publisher {
  using SharedPtr = std::shared_ptr<Publisher<M>>;
  std::string topic;
  void publish(const M &m) {
    auto &b = ref_shim::Board::get();
    b.last[topic] = std::make_shared<M>(m);
    b.count[topic]++;
  }
  void publish(std::unique_ptr<M> m) { publish(*m); }
  size_t get_subscription_count() const { return 1; }
};
template <class M>
struct Subscription { using SharedPtr = std::shared_ptr<Subscription<M>>; };
template <class S>
struct Service { using SharedPtr = std::shared_ptr<Service<S>>; };
template <class S>
struct Client {
  using SharedPtr = std::shared_ptr<Client<S>>;
  using SharedFuture = std::shared_future<typename S::Response::SharedPtr>;
  bool service_is_ready() const { return false; }
  template <class... A> bool wait_for_service(A &&...) { return false; }
  template <class R, class... A> SharedFuture async_send_request(R, A &&...) { return SharedFuture(); }
};
struct TimerBase { using SharedPtr = std::shared_ptr<TimerBase>; void cancel() {} };

struct Parameter {
  double num = 0;
  std::string str;
  bool is_str = false;
  double as_double() const { return num; }
  int64_t as_int() const { return (int64_t)num; }
  bool as_bool() const { return num != 0; }
  std::string as_string() const { return str; }
};

class Node {
 public:
  using SharedPtr = std::shared_ptr<Node>;
  explicit Node(const std::string &name) : name_(name), clock_(std::make_shared<Clock>()) {}
  virtual ~Node() = default;

  // declare_parameter<float>("x", -0.4): the default is converted to the declared type first (rclcpp does the same)
  template <class T = void, class D>
  void declare_parameter(const std::string &name, const D &def) {
    if constexpr (std::is_void<T>::value) set_default(name, def);
    else set_default(name, static_cast<T>(def));
  }

  Parameter get_parameter(const std::string &name) const {
    auto it = params_.find(name);
    return it == params_.end() ? Parameter() : it->second;
  }
  template <class T>
  bool get_parameter(const std::string &name, T &out) const {
    auto it = params_.find(name);
    if (it == params_.end()) return false;
    assign(out, it->second);
    return true;
  }

  template <class M, class... A>
  typename Publisher<M>::SharedPtr create_publisher(const std::string &topic, A &&...) {
    auto p = std::make_shared<Publisher<M>>();
    p->topic = topic;
    return p;
  }
  template <class M, class... A>
  typename Subscription<M>::SharedPtr create_subscription(const std::string &, A &&...) {
    return std::make_shared<Subscription<M>>();
  }
  template <class S, class... A>
  typename Service<S>::SharedPtr create_service(const std::string &, A &&...) { return std::make_shared<Service<S>>(); }
  template <class S, class... A>
  typename Client<S>::SharedPtr create_client(const std::string &, A &&...) { return std::make_shared<Client<S>>(); }
  template <class... A>
  TimerBase::SharedPtr create_wall_timer(A &&...) { return std::make_shared<TimerBase>(); }

  Clock::SharedPtr get_clock() const { return clock_; }
  Time now() const { return clock_->now(); }
  Logger get_logger() const { return Logger(); }
  const char *get_name() const { return name_.c_str(); }

 private:
  static void assign(std::string &o, const Parameter &p) { o = p.str; }
  static void assign(bool &o, const Parameter &p) { o = p.num != 0; }
  template <class T>
  static void assign(T &o, const Parameter &p) { o = static_cast<T>(p.num); }

  void set_default(const std::string &name, const std::string &v) {
    Parameter p; p.is_str = true; p.str = v;
    auto &ov = ref_shim::ParamOverrides::get().str;
    auto it = ov.find(name);
    if (it != ov.end()) p.str = it->second;
    params_[name] = p;
  }
  void set_default(const std::string &name, const char *v) { set_default(name, std::string(v)); }
  template <class T>
  void set_default(const std::string &name, T v) {
    Parameter p; p.num = static_cast<double>(v);   // declare_parameter<float>(.., -0.4): the default is rounded to T first
    auto &ov = ref_shim::ParamOverrides::get().num;
    auto it = ov.find(name);
    if (it != ov.end()) p.num = it->second;
    params_[name] = p;
  }
  std::string name_;
  Clock::SharedPtr clock_;
  std::map<std::string, Parameter> params_;
};

inline void init(int, char **) {}
inline void shutdown() {}
template <class T> inline void spin(T) {}
inline bool ok() { return true; }
}  // namespace rclcpp

#define RCLCPP_INFO(...) ((void)0)
#define RCLCPP_WARN(...) ((void)0)
#define RCLCPP_ERROR(...) ((void)0)
#define RCLCPP_DEBUG(...) ((void)0)
#define RCLCPP_INFO_THROTTLE(...) ((void)0)
#define RCLCPP_WARN_THROTTLE(...) ((void)0)
#define RCLCPP_ERROR_THROTTLE(...) ((void)0)
