#include <cstdio>
#include <fstream>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include "pe_codegen.h"
#include "ph_ron.h"
#include "ph_scene.h"
static std::string slurp(const char* p) { std::ifstream f(p, std::ios::binary); std::ostringstream s; s << f.rdbuf(); return s.str(); }
int main(int argc, char** argv) {
    std::mt19937 rng(777);
    long ok = 0, thrown = 0;
    for (int i = 1; i < argc; i++) {
        std::string err;
        ph::RonPtr root = ph::ron_parse(slurp(argv[i]), err);
        if (!root) continue;
        ph::Scene sc;
        if (!sc.load(*root)) continue;
        std::vector<std::string> snippets;
        for (auto& o : sc.objects) snippets.push_back(o.code);
        for (auto& m : sc.materials) if (!m.code.empty()) snippets.push_back(m.code);
        for (auto& l : sc.library) snippets.push_back(l.second);
        for (auto& l : sc.intersection_materials) snippets.push_back(l.second);
        for (const std::string& base : snippets) {
            for (int m = 0; m < 12; m++) {
                std::string t = base;
                if (m > 0) {
                    int n = 1 + int(rng() % 4);
                    for (int k = 0; k < n && !t.empty(); k++) {
                        size_t pos = rng() % t.size();
                        switch (rng() % 4) {
                            case 0: t.erase(pos, 1 + rng() % 20); break;
                            case 1: t.insert(pos, std::string(1, "(){}[];.,+-*/=<>!&|^%?:#\"'\\"[rng() % 27])); break;
                            case 2: t.insert(pos, (const char*[]){".xyzw", "/*", "*/", "//", "out ", "inout ", "1.e", "..", ".x.y", "#define A\n"}[rng() % 10]); break;
                            default: t.resize(pos); break;
                        }
                    }
                }
                std::set<std::string> swz, lsw;
                try { std::string o = pe_host::glsl_to_cuda(t, swz, (m & 1) != 0, &lsw); ok++; (void)o; }
                catch (const std::exception&) { thrown++; }
            }
        }
    }
    std::printf("rewritten %ld rejected %ld\n", ok, thrown);
    return 0;
}
