// Stand-in for the handful of MOOSE framework symbols Marlin's host objects use:
// InputParameters, MooseEnum, MooseObject, registerMooseObject / Factory, mooseError / paramError,
// DependencyResolverInterface::sort, ExecFlagType.  MOOSE itself cannot be compiled in this image
// (libMesh, PETSc and WASP are absent), so the host classes in host/src are written against this
// header; the spellings are MOOSE's so that the same class bodies compile against the real framework.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

using Real = double;
// MOOSE's strong string typedefs are plain strings here
using TensorInputBufferName = std::string;
using TensorOutputBufferName = std::string;
using TensorComputeName = std::string;
using FunctionName = std::string;
using MarlinConstantName = std::string;
using PostprocessorName = std::string;

struct MooseException : std::runtime_error {
  using std::runtime_error::runtime_error;
};

template <typename... A>
std::string mooseStringify(const A &...a) {
  std::ostringstream os;
  os.precision(17);
  (os << ... << a);
  return os.str();
}
template <typename... A>
[[noreturn]] void mooseError(const A &...a) {
  throw MooseException(mooseStringify(a...));
}
template <typename... A>
void mooseWarning(const A &...a) {
  std::cerr << "*** Warning ***\n" << mooseStringify(a...) << "\n";
}
template <typename... A>
void mooseInfo(const A &...a) {
  std::cerr << "*** Info ***\n" << mooseStringify(a...) << "\n";
}

// ---- MooseEnum ------------------------------------------------------------------------------
// "X=0 Y=1 Z=2" or "REAL RECIPROCAL NONE"; comparison is case-insensitive like MOOSE's.
class MooseEnum {
public:
  MooseEnum() = default;
  MooseEnum(const std::string &names, const std::string &dflt = "");
  MooseEnum &operator=(const std::string &v);
  bool isValid() const { return _cur >= 0; }
  operator int() const { return _cur < 0 ? -1 : _ids[_cur]; }
  operator std::string() const { return _cur < 0 ? "" : _names[_cur]; }
  bool operator==(const char *s) const;
  template <typename E>
  E getEnum() const {
    return static_cast<E>(int(*this));
  }
  const std::vector<std::string> &names() const { return _names; }
  std::string raw() const;  // "A B C" form for stringification

private:
  std::vector<std::string> _names;
  std::vector<int> _ids;
  int _cur = -1;
};
using MultiMooseEnum = std::vector<std::string>;

// ---- InputParameters ------------------------------------------------------------------------
// Values are kept as the input-file text and converted on getParam<T>(); defaults are stored the
// same way.  Pointer-valued private parameters (_tensor_problem, _domain) live in a side table.
namespace shim_detail {
std::vector<std::string> splitList(const std::string &s);
template <typename T>
struct Conv;
template <>
struct Conv<std::string> {
  static std::string from(const std::string &s, const std::string &) { return s; }
  static std::string to(const std::string &v) { return v; }
};
template <>
struct Conv<double> {
  static double from(const std::string &s, const std::string &what);
  static std::string to(double v) { return mooseStringify(v); }
};
template <>
struct Conv<bool> {
  static bool from(const std::string &s, const std::string &what);
  static std::string to(bool v) { return v ? "true" : "false"; }
};
template <typename I>
struct ConvInt {
  static I from(const std::string &s, const std::string &what) {
    const double d = Conv<double>::from(s, what);
    if (d != static_cast<double>(static_cast<long long>(d))) mooseError(what, ": '", s, "' is not an integer");
    if (std::is_unsigned<I>::value && d < 0) mooseError(what, ": '", s, "' must not be negative");
    return static_cast<I>(static_cast<long long>(d));
  }
  static std::string to(I v) { return std::to_string(v); }
};
template <> struct Conv<int> : ConvInt<int> {};
template <> struct Conv<unsigned int> : ConvInt<unsigned int> {};
template <> struct Conv<long> : ConvInt<long> {};
template <> struct Conv<unsigned long> : ConvInt<unsigned long> {};
template <> struct Conv<long long> : ConvInt<long long> {};
template <typename T>
struct Conv<std::vector<T>> {
  static std::vector<T> from(const std::string &s, const std::string &what) {
    std::vector<T> out;
    for (const auto &w : splitList(s)) out.push_back(Conv<T>::from(w, what));
    return out;
  }
  static std::string to(const std::vector<T> &v) {
    std::string s;
    for (const auto &e : v) s += (s.empty() ? "" : " ") + Conv<T>::to(e);
    return s;
  }
};
}  // namespace shim_detail

class InputParameters {
public:
  struct Entry {
    std::string value, doc, range;
    bool set = false, required = false, is_private = false, user_set = false;
    MooseEnum enum_proto;
    bool is_enum = false;
  };

  template <typename T>
  void addParam(const std::string &name, const std::string &doc) {
    auto &e = _entries[name];
    e.doc = doc;
  }
  template <typename T>
  void addParam(const std::string &name, const T &dflt, const std::string &doc) {
    auto &e = _entries[name];
    e.doc = doc;
    setDefault<T>(e, dflt);
  }
  template <typename T>
  void addRequiredParam(const std::string &name, const std::string &doc) {
    auto &e = _entries[name];
    e.doc = doc;
    e.required = true;
  }
  template <typename T>
  void addRequiredParam(const std::string &name, const T &proto, const std::string &doc) {  // MooseEnum form
    auto &e = _entries[name];
    e.doc = doc;
    e.required = true;
    setDefault<T>(e, proto);
    e.set = false;
  }
  template <typename T>
  void addRangeCheckedParam(const std::string &name, const T &dflt, const std::string &range, const std::string &doc) {
    addParam<T>(name, dflt, doc);
    _entries[name].range = range;
  }
  template <typename T>
  void addPrivateParam(const std::string &name, T) {
    _entries[name].is_private = true;
  }
  template <typename T>
  void suppressParameter(const std::string &name) {
    _entries[name].is_private = true;
  }
  void addClassDescription(const std::string &d) { _class_description = d; }
  void registerBase(const std::string &b) { _base = b; }
  const std::string &base() const { return _base; }
  const std::string &classDescription() const { return _class_description; }

  bool have(const std::string &name) const { return _entries.count(name) != 0; }
  bool isParamValid(const std::string &name) const {
    auto it = _entries.find(name);
    return it != _entries.end() && it->second.set;
  }
  bool isParamSetByUser(const std::string &name) const {
    auto it = _entries.find(name);
    return it != _entries.end() && it->second.user_set;
  }
  // text from the input file
  void setFromInput(const std::string &name, const std::string &text);
  template <typename T>
  void set(const std::string &name, const T &v) {
    auto &e = _entries[name];
    e.value = shim_detail::Conv<T>::to(v);
    e.set = true;
  }
  template <typename T>
  T get(const std::string &name, const std::string &object_path) const;
  void setPointer(const std::string &name, void *p) { _pointers[name] = p; }
  void *pointer(const std::string &name) const {
    auto it = _pointers.find(name);
    return it == _pointers.end() ? nullptr : it->second;
  }
  const std::map<std::string, Entry> &entries() const { return _entries; }
  // required parameters present, range checks (throws with the object path)
  void check(const std::string &object_path) const;
  InputParameters &operator+=(const InputParameters &o);

private:
  template <typename T>
  void setDefault(Entry &e, const T &dflt) {
    if constexpr (std::is_same<T, MooseEnum>::value) {
      e.enum_proto = dflt;
      e.is_enum = true;
      e.value = std::string(e.enum_proto);
      e.set = e.enum_proto.isValid();
    } else {
      e.value = shim_detail::Conv<T>::to(dflt);
      e.set = true;
    }
  }
  std::map<std::string, Entry> _entries;
  std::map<std::string, void *> _pointers;
  std::string _class_description, _base;
};

template <typename T>
T InputParameters::get(const std::string &name, const std::string &object_path) const {
  auto it = _entries.find(name);
  if (it == _entries.end()) mooseError(object_path, ": parameter '", name, "' is not declared");
  if (!it->second.set) mooseError(object_path, "/", name, ": missing required parameter");
  if constexpr (std::is_same<T, MooseEnum>::value) {
    MooseEnum e = it->second.enum_proto;
    try {
      e = it->second.value;
    } catch (const MooseException &x) {
      mooseError(object_path, "/", name, ": ", x.what());
    }
    return e;
  } else {
    return shim_detail::Conv<T>::from(it->second.value, object_path + "/" + name);
  }
}

InputParameters emptyInputParameters();

// ---- MooseObject ------------------------------------------------------------------------------
class MooseObject {
public:
  explicit MooseObject(const InputParameters &p) : _pars(p), _name(p.get<std::string>("_object_name", "?")), _type(p.get<std::string>("_type", "?")), _path(p.get<std::string>("_object_path", "?")) {}
  virtual ~MooseObject() = default;
  static InputParameters validParams();
  const std::string &name() const { return _name; }
  const std::string &type() const { return _type; }
  const InputParameters &parameters() const { return _pars; }
  template <typename T>
  T getParam(const std::string &n) const {
    return _pars.get<T>(n, _path);
  }
  // pairs two equally long vector parameters (MooseObject::getParam<T1,T2>(p1, p2))
  template <typename T1, typename T2>
  std::vector<std::pair<T1, T2>> getParam(const std::string &n1, const std::string &n2) const {
    std::vector<T1> a = isParamValid(n1) ? getParam<std::vector<T1>>(n1) : std::vector<T1>();
    std::vector<T2> b = isParamValid(n2) ? getParam<std::vector<T2>>(n2) : std::vector<T2>();
    if (a.size() != b.size()) paramError(n1, "Vector parameters '", n1, "' and '", n2, "' must have the same length.");
    std::vector<std::pair<T1, T2>> out;
    for (size_t i = 0; i < a.size(); ++i) out.emplace_back(a[i], b[i]);
    return out;
  }
  bool isParamValid(const std::string &n) const { return _pars.isParamValid(n); }
  bool isParamSetByUser(const std::string &n) const { return _pars.isParamSetByUser(n); }
  template <typename T>
  T *getCheckedPointerParam(const std::string &n) const {
    void *p = _pars.pointer(n);
    if (!p) mooseError(_path, ": internal pointer parameter '", n, "' is not set");
    return static_cast<T *>(p);
  }
  template <typename... A>
  [[noreturn]] void mooseError(const A &...a) const {
    ::mooseError("The following error occurred in the ", _pars.base().empty() ? "object" : _pars.base(), " '", _name, "' of type ", _type, ".\n\n", a...);
  }
  template <typename... A>
  [[noreturn]] void paramError(const std::string &param, const A &...a) const {
    ::mooseError(_path, "/", param, ": ", a...);
  }
  template <typename... A>
  void mooseWarning(const A &...a) const {
    ::mooseWarning(_path, ": ", a...);
  }

protected:
  const InputParameters _pars;
  const std::string _name, _type, _path;
};

// ---- Factory / registry ----------------------------------------------------------------------
class Factory {
public:
  using Build = std::function<std::shared_ptr<MooseObject>(const InputParameters &)>;
  using Params = std::function<InputParameters()>;
  struct Item {
    Build build;
    Params params;
    std::string app, file;
  };
  static Factory &instance();
  void reg(const std::string &name, Item item) { _items[name] = std::move(item); }
  bool isRegistered(const std::string &name) const { return _items.count(name) != 0; }
  InputParameters getValidParams(const std::string &name) const;
  std::shared_ptr<MooseObject> create(const std::string &type, const InputParameters &p) const;
  std::vector<std::string> registeredNames() const;

private:
  std::map<std::string, Item> _items;
};

template <typename T>
struct FactoryRegistrar {
  FactoryRegistrar(const char *app, const char *name, const char *file) {
    Factory::instance().reg(name, Factory::Item{[](const InputParameters &p) -> std::shared_ptr<MooseObject> { return std::make_shared<T>(p); },
                                                []() { return T::validParams(); }, app, file});
  }
};
#define SHIM_CAT2(a, b) a##b
#define SHIM_CAT(a, b) SHIM_CAT2(a, b)
#define registerMooseObject(app, Class) static FactoryRegistrar<Class> SHIM_CAT(shim_registrar_, __COUNTER__)(app, #Class, __FILE__)
#define registerMooseObjectAliased(app, Class, alias) static FactoryRegistrar<Class> SHIM_CAT(shim_registrar_, __COUNTER__)(app, alias, __FILE__)
#define registerMooseObjectRenamed(app, OldName, date, Class) static FactoryRegistrar<Class> SHIM_CAT(shim_registrar_, __COUNTER__)(app, #OldName, __FILE__)

// ---- DependencyResolverInterface ---------------------------------------------------------------
class DependencyResolverInterface {
public:
  virtual ~DependencyResolverInterface() = default;
  virtual const std::set<std::string> &getRequestedItems() = 0;
  virtual const std::set<std::string> &getSuppliedItems() = 0;
  // Stable topological sort: an object runs after every object that supplies one of the items it
  // requests; objects not ordered by a dependency keep their input-file order.  Throws on a cycle.
  template <typename T>
  static void sort(std::vector<std::shared_ptr<T>> &v) {
    const size_t n = v.size();
    std::vector<std::vector<size_t>> succ(n);
    std::vector<int> indeg(n, 0);
    for (size_t a = 0; a < n; ++a)
      for (size_t b = 0; b < n; ++b) {
        if (a == b) continue;
        bool dep = false;  // b depends on a
        for (const auto &item : v[b]->getRequestedItems())
          if (v[a]->getSuppliedItems().count(item)) dep = true;
        if (dep) {
          succ[a].push_back(b);
          ++indeg[b];
        }
      }
    std::vector<std::shared_ptr<T>> out;
    std::vector<bool> done(n, false);
    for (size_t k = 0; k < n; ++k) {
      size_t pick = n;
      for (size_t i = 0; i < n; ++i)
        if (!done[i] && indeg[i] == 0) {
          pick = i;
          break;
        }
      if (pick == n) {
        std::string names;
        for (size_t i = 0; i < n; ++i)
          if (!done[i]) names += " " + v[i]->name();
        ::mooseError("Cyclic dependency detected between the objects:", names);
      }
      done[pick] = true;
      out.push_back(v[pick]);
      for (size_t s : succ[pick]) --indeg[s];
    }
    v.swap(out);
  }
};

// ---- execute_on flags ------------------------------------------------------------------------
enum ExecFlagType { EXEC_NONE = 0, EXEC_INITIAL = 1, EXEC_TIMESTEP_BEGIN = 2, EXEC_TIMESTEP_END = 4, EXEC_FINAL = 8 };
int parseExecFlags(const std::string &text, const std::string &what);
