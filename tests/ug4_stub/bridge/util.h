// MINIMAL STAND-IN for UG4's ugcore `bridge/util.h` (ug::bridge::Registry, SmartPtr, UG_THROW) -- test infrastructure.
// No UG4 tree exists in this image, so plugins/ADMMOptimB200/admm_b200_plugin.cpp is compiled and its
// registrations are exercised against this stub: it records every class, constructor, method and function name a plugin registers
// (tests/test_host.py::test_ug4_plugin_shim_registers_the_names_the_scripts_call).  The member templates mirror the call shapes of
// ugcore's registry [UPSTREAM-UNVERIFIED]: reg.add_class_<T>(name, grp).add_constructor().template add_constructor<Sig>(doc)
// .add_method(name, &T::m, ...).set_construct_as_smart_pointer(true); reg.add_function(name, &f, grp, ...).
#pragma once
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

typedef double number;

template <typename T>
using SmartPtr = std::shared_ptr<T>;
template <typename T>
using ConstSmartPtr = std::shared_ptr<const T>;
template <typename T, typename... A>
SmartPtr<T> make_sp_new(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

namespace ug {
struct UGError : std::runtime_error {
    explicit UGError(const std::string& m) : std::runtime_error(m) {}
};
namespace bridge {

struct ExportedClassInfo {
    std::string name, group;
    int constructors = 0;
    std::vector<std::string> methods;
};

template <typename T>
class ExportedClass {
 public:
    explicit ExportedClass(ExportedClassInfo* i) : info_(i) {}
    ExportedClass& add_constructor() { info_->constructors++; return *this; }
    template <typename Sig>
    ExportedClass& add_constructor(const std::string& = "", const std::string& = "", const std::string& = "") { info_->constructors++; return *this; }
    template <typename M>
    ExportedClass& add_method(const std::string& name, M, const std::string& = "", const std::string& = "", const std::string& = "", const std::string& = "") {
        info_->methods.push_back(name);
        return *this;
    }
    ExportedClass& set_construct_as_smart_pointer(bool) { return *this; }
 private:
    ExportedClassInfo* info_;
};

class Registry {
 public:
    template <typename T>
    ExportedClass<T> add_class_(const std::string& name, const std::string& grp = "", const std::string& = "") {
        classes_.push_back(std::unique_ptr<ExportedClassInfo>(new ExportedClassInfo()));
        classes_.back()->name = name;
        classes_.back()->group = grp;
        return ExportedClass<T>(classes_.back().get());
    }
    template <typename T, typename TBase>
    ExportedClass<T> add_class_(const std::string& name, const std::string& grp = "", const std::string& tt = "") {
        static_assert(std::is_base_of<TBase, T>::value, "registered base class is not a base");
        return add_class_<T>(name, grp, tt);
    }
    template <typename F>
    Registry& add_function(const std::string& name, F, const std::string& grp = "", const std::string& = "", const std::string& = "", const std::string& = "") {
        functions_.push_back(name);
        return *this;
    }
    void add_class_to_group(const std::string& cls, const std::string& group, const std::string& = "") { groups_[group].push_back(cls); }
    const std::vector<std::unique_ptr<ExportedClassInfo>>& classes() const { return classes_; }
    const std::vector<std::string>& functions() const { return functions_; }
 private:
    std::vector<std::unique_ptr<ExportedClassInfo>> classes_;
    std::vector<std::string> functions_;
    std::map<std::string, std::vector<std::string>> groups_;
};

}  // namespace bridge
}  // namespace ug

#define UG_THROW(msg)                         \
    do {                                      \
        std::stringstream ss__;               \
        ss__ << msg;                          \
        throw ug::UGError(ss__.str());        \
    } while (0)
#define UG_REGISTRY_CATCH_THROW(grp) \
    catch (const ug::UGError& e) { throw; }
