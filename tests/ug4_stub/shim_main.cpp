// Test driver of the UG4 plugin shim: load the plugin's entry point against the stand-in registry and print what it registered
// (tests/test_host.py::test_ug4_plugin_shim_registers_the_names_the_scripts_call).  Registration creates no GPU context.
#include <cstdio>

#include "bridge/util.h"

extern "C" void InitUGPlugin_ADMMOptimB200(ug::bridge::Registry* reg, std::string grp);

int main() {
    ug::bridge::Registry reg;
    InitUGPlugin_ADMMOptimB200(&reg, "ug4");
    std::printf("{\"classes\": {");
    bool first = true;
    for (const auto& c : reg.classes()) {
        std::printf("%s\"%s\": {\"group\": \"%s\", \"constructors\": %d, \"methods\": [", first ? "" : ", ", c->name.c_str(), c->group.c_str(), c->constructors);
        for (size_t i = 0; i < c->methods.size(); ++i) std::printf("%s\"%s\"", i ? ", " : "", c->methods[i].c_str());
        std::printf("]}");
        first = false;
    }
    std::printf("}, \"functions\": [");
    for (size_t i = 0; i < reg.functions().size(); ++i) std::printf("%s\"%s\"", i ? ", " : "", reg.functions()[i].c_str());
    std::printf("]}\n");
    return 0;
}
