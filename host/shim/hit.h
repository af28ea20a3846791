// Minimal reader for the subset of MOOSE's "hit" input syntax that Marlin inputs use
// (SURVEY.md 8b): [block] / [] sections (also the legacy [./name] / [../] form), `key = value`
// fields with bare, 'single' or "double" quoted (possibly multi-line) values, # comments,
// top-level and scoped variables with ${name} substitution, ${fparse expr}, ${units value ...},
// ${raw a b}, `active = '...'` / `inactive = '...'` filtering, and command-line overrides
// (`Block/sub/key=value`, `key=value`).  MOOSE's own parser cannot be linked here (no libMesh /
// WASP in this image), so this stands in for it; the syntax handled is MOOSE's, not a new one.
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace hit {

struct Node {
  bool is_section = true;
  std::string name;        // section name or field key
  std::string value;       // field value (quotes removed, ${...} expanded)
  bool quoted = false;
  int line = 0;
  Node *parent = nullptr;
  std::vector<std::unique_ptr<Node>> children;

  std::string fullpath() const;
  Node *find(const std::string &path);                 // "a/b/c", sections or fields
  const Node *find(const std::string &path) const;
  std::vector<Node *> sections() const;                // child sections honouring active / inactive
  std::vector<Node *> fields() const;
  const Node *field(const std::string &key) const;     // direct child field or nullptr
};

// Parses `text`; applies `overrides` ("path/key=value"); expands ${...}.  Throws std::runtime_error
// with "file:line: message" on malformed input.
std::unique_ptr<Node> parse(const std::string &text, const std::string &fname, const std::vector<std::string> &overrides);

}  // namespace hit
