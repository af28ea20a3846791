#include "moose_shim.h"

#include <cctype>
#include <cstdlib>

namespace {
std::string upper(std::string s) {
  for (auto &c : s) c = std::toupper((unsigned char)c);
  return s;
}
}  // namespace

// ---- MooseEnum --------------------------------------------------------------------------------
MooseEnum::MooseEnum(const std::string &names, const std::string &dflt) {
  int next = 0;
  for (const auto &w : shim_detail::splitList(names)) {
    const size_t eq = w.find('=');
    std::string n = w.substr(0, eq);
    int id = next;
    if (eq != std::string::npos) id = std::atoi(w.substr(eq + 1).c_str());
    _names.push_back(n);
    _ids.push_back(id);
    next = id + 1;
  }
  if (!dflt.empty()) *this = dflt;
}
MooseEnum &MooseEnum::operator=(const std::string &v) {
  const std::string u = upper(v);
  for (size_t i = 0; i < _names.size(); ++i)
    if (upper(_names[i]) == u) {
      _cur = (int)i;
      return *this;
    }
  mooseError("Invalid option \"", v, "\" in MooseEnum.  Valid options (not case-sensitive) are \"", raw(), "\".");
}
bool MooseEnum::operator==(const char *s) const { return _cur >= 0 && upper(_names[_cur]) == upper(s); }
std::string MooseEnum::raw() const {
  std::string s;
  for (const auto &n : _names) s += (s.empty() ? "" : " ") + n;
  return s;
}

// ---- conversions --------------------------------------------------------------------------------
namespace shim_detail {
std::vector<std::string> splitList(const std::string &s) {
  std::vector<std::string> out;
  std::string w;
  for (char c : s) {
    if (std::isspace((unsigned char)c) || c == ';') {  // MOOSE accepts whitespace separated lists
      if (!w.empty()) out.push_back(w);
      w.clear();
    } else {
      w += c;
    }
  }
  if (!w.empty()) out.push_back(w);
  return out;
}
double Conv<double>::from(const std::string &s, const std::string &what) {
  const char *b = s.c_str();
  char *e = nullptr;
  const double v = std::strtod(b, &e);
  while (e && *e && std::isspace((unsigned char)*e)) ++e;
  if (e == b || (e && *e)) mooseError(what, ": invalid number '", s, "'");
  return v;
}
bool Conv<bool>::from(const std::string &s, const std::string &what) {
  const std::string u = upper(s);
  if (u == "TRUE" || u == "1" || u == "ON" || u == "YES") return true;
  if (u == "FALSE" || u == "0" || u == "OFF" || u == "NO") return false;
  mooseError(what, ": invalid boolean '", s, "'");
}
}  // namespace shim_detail

// ---- InputParameters ------------------------------------------------------------------------------
void InputParameters::setFromInput(const std::string &name, const std::string &text) {
  auto &e = _entries[name];
  e.value = text;
  e.set = true;
  e.user_set = true;
}

void InputParameters::check(const std::string &object_path) const {
  for (const auto &[name, e] : _entries) {
    if (e.required && !e.set) mooseError(object_path, "/", name, ": missing required parameter '", name, "'\n\tDoc String: \"", e.doc, "\"");
    if (e.is_enum && e.set) {
      MooseEnum probe = e.enum_proto;
      try {
        probe = e.value;
      } catch (const MooseException &x) {
        mooseError(object_path, "/", name, ": ", x.what());
      }
    }
    if (!e.range.empty() && e.set) {
      // range expressions are conjunctions of `<name> <op> <number>` joined by '&'
      const double v = shim_detail::Conv<double>::from(e.value, object_path + "/" + name);
      std::string r = e.range;
      size_t pos = 0;
      bool ok = true;
      while (pos < r.size()) {
        size_t amp = r.find('&', pos);
        std::string clause = r.substr(pos, amp == std::string::npos ? std::string::npos : amp - pos);
        pos = amp == std::string::npos ? r.size() : amp + 1;
        size_t k = clause.find_first_of("<>=!");
        if (k == std::string::npos) continue;
        size_t k2 = k;
        while (k2 < clause.size() && std::string("<>=!").find(clause[k2]) != std::string::npos) ++k2;
        const std::string op = clause.substr(k, k2 - k);
        const double rhs = std::strtod(clause.c_str() + k2, nullptr);
        if (op == "<") ok = ok && v < rhs;
        else if (op == "<=") ok = ok && v <= rhs;
        else if (op == ">") ok = ok && v > rhs;
        else if (op == ">=") ok = ok && v >= rhs;
        else if (op == "=" || op == "==") ok = ok && v == rhs;
        else if (op == "!=") ok = ok && v != rhs;
      }
      if (!ok) mooseError(object_path, "/", name, ": Range check failed; expression = '", e.range, "', value = ", e.value);
    }
  }
}

InputParameters &InputParameters::operator+=(const InputParameters &o) {
  for (const auto &kv : o._entries) _entries[kv.first] = kv.second;
  for (const auto &kv : o._pointers) _pointers[kv.first] = kv.second;
  if (!o._class_description.empty()) _class_description = o._class_description;
  if (!o._base.empty()) _base = o._base;
  return *this;
}

InputParameters emptyInputParameters() { return InputParameters(); }

InputParameters MooseObject::validParams() {
  InputParameters p;
  p.addPrivateParam<std::string>("_object_name", "");
  p.addPrivateParam<std::string>("_type", "");
  p.addPrivateParam<std::string>("_object_path", "");
  p.addParam<std::string>("type", "The object type");
  p.addParam<bool>("enable", true, "Set the enabled status of the MooseObject.");
  p.addParam<std::vector<std::string>>("control_tags", "Control tags (unused here)");
  return p;
}

// ---- Factory ----------------------------------------------------------------------------------------
Factory &Factory::instance() {
  static Factory f;
  return f;
}
InputParameters Factory::getValidParams(const std::string &name) const {
  auto it = _items.find(name);
  if (it == _items.end()) mooseError("A '", name, "' is not a registered object.");
  return it->second.params();
}
std::shared_ptr<MooseObject> Factory::create(const std::string &type, const InputParameters &p) const {
  auto it = _items.find(type);
  if (it == _items.end()) mooseError("A '", type, "' is not a registered object.");
  return it->second.build(p);
}
std::vector<std::string> Factory::registeredNames() const {
  std::vector<std::string> out;
  for (const auto &kv : _items) out.push_back(kv.first);
  return out;
}

int parseExecFlags(const std::string &text, const std::string &what) {
  int flags = 0;
  for (const auto &w : shim_detail::splitList(text)) {
    const std::string u = upper(w);
    if (u == "INITIAL") flags |= EXEC_INITIAL;
    else if (u == "TIMESTEP_BEGIN") flags |= EXEC_TIMESTEP_BEGIN;
    else if (u == "TIMESTEP_END") flags |= EXEC_TIMESTEP_END;
    else if (u == "FINAL") flags |= EXEC_FINAL;
    else if (u == "NONE") flags |= 0;
    else if (u == "LINEAR" || u == "NONLINEAR" || u == "ALWAYS" || u == "CUSTOM" || u == "SUBDOMAIN" || u == "FAILED") flags |= 0;
    else mooseError(what, ": unknown execute_on flag '", w, "'");
  }
  return flags;
}
