// STAND-IN for the slice of yaml-cpp that graph_core touches (Node::toYAML / fromYAML, Tree/Path
// round trips).  Test infrastructure for oracle/_ref only.  Handle semantics like yaml-cpp: copies of
// a Node alias the same underlying value.
#pragma once
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
namespace YAML
{
namespace EmitterStyle
{
enum value { Default, Block, Flow };
}
class Node
{
  struct Data
  {
    enum Kind { Null, Scalar, Sequence, Map } kind = Null;
    std::string scalar;
    std::vector<Node> seq;
    std::vector<std::pair<std::string, Node>> map;
    bool flow = false;
    bool defined = true;
  };
  std::shared_ptr<Data> d_;

public:
  Node() : d_(std::make_shared<Data>()) {}
  template <typename T>
  Node(const T& v) : d_(std::make_shared<Data>()) { *this = v; }
  Node(const Node&) = default;
  Node& operator=(const Node& o)
  {
    if (d_ && !d_->defined && o.d_)  // assignment through operator[] on a missing key materialises it
    {
      *d_ = *o.d_;
      d_->defined = true;
    }
    else
      d_ = o.d_;
    return *this;
  }
  template <typename T>
  Node& operator=(const T& v)
  {
    std::ostringstream ss;
    ss.precision(17);
    ss << v;
    d_->kind = Data::Scalar;
    d_->scalar = ss.str();
    d_->defined = true;
    return *this;
  }
  void SetStyle(EmitterStyle::value s) { d_->flow = (s == EmitterStyle::Flow); }
  bool IsDefined() const { return d_->defined; }
  bool IsNull() const { return d_->kind == Data::Null; }
  bool IsScalar() const { return d_->kind == Data::Scalar; }
  bool IsSequence() const { return d_->kind == Data::Sequence; }
  bool IsMap() const { return d_->kind == Data::Map; }
  explicit operator bool() const { return d_->defined; }
  bool operator!() const { return !d_->defined; }
  std::size_t size() const { return d_->kind == Data::Sequence ? d_->seq.size() : d_->map.size(); }
  void push_back(const Node& n)
  {
    d_->kind = Data::Sequence;
    d_->defined = true;
    d_->seq.push_back(n);
  }
  template <typename T>
  void push_back(const T& v) { push_back(Node(v)); }
  Node operator[](std::size_t i) const { return d_->seq.at(i); }
  Node operator[](int i) const { return d_->seq.at((std::size_t)i); }
  // non-const twins: without them `node[0]` on a non-const Node is ambiguous with the const char* key overload
  Node operator[](std::size_t i) { return d_->seq.at(i); }
  Node operator[](int i) { return d_->seq.at((std::size_t)i); }
  Node operator[](const std::string& k) const
  {
    for (auto& kv : d_->map)
      if (kv.first == k)
        return kv.second;
    Node missing;
    missing.d_->defined = false;
    return missing;
  }
  Node operator[](const char* k) const { return (*this)[std::string(k)]; }
  Node operator[](const std::string& k)
  {
    for (auto& kv : d_->map)
      if (kv.first == k)
        return kv.second;
    d_->kind = Data::Map;
    d_->defined = true;
    Node fresh;
    fresh.d_->defined = false;  // materialised by the assignment that follows (see operator=)
    d_->map.emplace_back(k, fresh);
    return d_->map.back().second;
  }
  Node operator[](const char* k) { return (*this)[std::string(k)]; }
  template <typename T>
  T as() const
  {
    if (d_->kind != Data::Scalar)
      throw std::runtime_error("YAML stand-in: as<T>() on a non-scalar");
    std::istringstream ss(d_->scalar);
    T v{};
    if (d_->scalar == "true" || d_->scalar == "false")
    {
      std::istringstream sb(d_->scalar);
      sb >> std::boolalpha >> v;
      return v;
    }
    ss >> v;
    return v;
  }
  std::vector<Node>::const_iterator begin() const { return d_->seq.begin(); }
  std::vector<Node>::const_iterator end() const { return d_->seq.end(); }
  void emit(std::ostream& os, int indent, bool flow) const
  {
    flow = flow || d_->flow;
    switch (d_->kind)
    {
      case Data::Null:
        os << "~";
        break;
      case Data::Scalar:
        os << d_->scalar;
        break;
      case Data::Sequence:
        if (flow)
        {
          os << "[";
          for (std::size_t i = 0; i < d_->seq.size(); i++)
          {
            if (i)
              os << ", ";
            d_->seq[i].emit(os, 0, true);
          }
          os << "]";
        }
        else
          for (std::size_t i = 0; i < d_->seq.size(); i++)
          {
            if (i)
              os << "\n" << std::string((std::size_t)indent, ' ');
            os << "- ";
            d_->seq[i].emit(os, indent + 2, false);
          }
        break;
      case Data::Map:
        for (std::size_t i = 0; i < d_->map.size(); i++)
        {
          if (i)
            os << "\n" << std::string((std::size_t)indent, ' ');
          os << d_->map[i].first << ":";
          const Node& c = d_->map[i].second;
          bool inl = c.d_->kind == Data::Scalar || c.d_->kind == Data::Null || c.d_->flow;
          if (inl)
          {
            os << " ";
            c.emit(os, indent + 2, false);
          }
          else
          {
            os << "\n" << std::string((std::size_t)indent + 2, ' ');
            c.emit(os, indent + 2, false);
          }
        }
        break;
    }
  }
};
template <>
inline Node& Node::operator=<bool>(const bool& v)
{
  d_->kind = Data::Scalar;
  d_->scalar = v ? "true" : "false";
  d_->defined = true;
  return *this;
}
inline std::ostream& operator<<(std::ostream& os, const Node& n)
{
  n.emit(os, 0, false);
  return os;
}
}  // namespace YAML
