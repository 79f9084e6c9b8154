// Test driver of the UG4 plugin shim: load the plugin's entry point against the stand-in registry and print what it registered
// (tests/test_host.py::test_ug4_plugin_shim_registers_the_names_the_scripts_call).  Registration creates no GPU context.
//   shim_test                                   -> JSON of the registered classes / methods / functions
//   shim_test ugx <grid.ugx> <refs> <level> <out.ugx>   -> the shim's SaveGridLevelToFile on a host-only domain (no GPU)
//   shim_test vtu <out.vtu>                             -> the shim's .vtu writer on a small synthetic 3D function
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bridge/util.h"

#include "admm_b200.h"

extern "C" void InitUGPlugin_ADMMOptimB200(ug::bridge::Registry* reg, std::string grp);
namespace ug {
namespace ADMMOptimB200 {
void SaveGridLevelToFileB200(ab_domain* dom, int level, const char* filename);
void WriteVTUB200(const char* path, int dim, int nv, const double* xyz, int ne, const int32_t* elems, const std::vector<std::string>& names,
                  const std::vector<std::vector<int>>& comps, int nfct, const double* values);
}  // namespace ADMMOptimB200
}  // namespace ug

int main(int argc, char** argv) {
    if (argc == 6 && std::strcmp(argv[1], "ugx") == 0) {
        ab_domain* dom = nullptr;
        if (ab_domain_load_ugx(nullptr, argv[2], &dom) != AB_OK || ab_domain_refine(dom, std::atoi(argv[3])) != AB_OK) {
            std::fprintf(stderr, "%s\n", ab_last_error());
            return 1;
        }
        ug::ADMMOptimB200::SaveGridLevelToFileB200(dom, std::atoi(argv[4]), argv[5]);
        ab_domain_destroy(dom);
        return 0;
    }
    if (argc == 3 && std::strcmp(argv[1], "vtu") == 0) {
        const double xyz[15] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1};
        const int32_t el[8] = {0, 1, 2, 3, 1, 2, 3, 4};
        double vals[15];
        for (int i = 0; i < 15; ++i) vals[i] = 0.1 * i - 0.3;
        ug::ADMMOptimB200::WriteVTUB200(argv[2], 3, 5, xyz, 2, el, {"u", "first"}, {{0, 1, 2}, {0}}, 3, vals);
        return 0;
    }
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
