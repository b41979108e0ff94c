// Formula evaluator of the scene front-end.
//
// The reference evaluates `Formula` / `FormulaInt` uniforms with the `fasteval` crate 0.2.4
// (/root/reference/src/gui/uniform.rs:602-635 parse + compile, :1009-1140 evaluation with a
// callback that supplies custom functions and the other uniforms as free variables).  The crate
// is not vendored; this is an implementation of its documented expression language:
//   precedence, tightest first:  ^  %  /  *  -  +  comparisons  and  or  (each on its own level);
//   ^ right-associative; a/b/c evaluated as a*(1/b)*(1/c); a-b-c as a+(-b)+(-c); literal
//   constants of a sum / product folded together and applied last; unary + - ! bind to the next
//   value; built-ins int ceil floor abs sign log round min max e pi sin cos tan asin acos atan sinh
//   cosh tanh asinh acosh atanh; every other name -- with or without arguments -- is resolved by
//   the caller's namespace callback.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace ph {

struct FormulaNode;
using FormulaPtr = std::shared_ptr<FormulaNode>;

// name, evaluated arguments -> value; return false if the name is unknown (evaluation fails)
using FormulaNamespace = std::function<bool(const std::string&, const std::vector<double>&, double&)>;

// Returns nullptr and sets err on a syntax error.
FormulaPtr formula_parse(const std::string& text, std::string& err);
// Returns false (and sets err) when a name cannot be resolved.
bool formula_eval(const FormulaPtr& f, const FormulaNamespace& ns, double& out, std::string& err);

}  // namespace ph
