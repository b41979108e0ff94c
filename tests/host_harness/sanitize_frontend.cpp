#include <cstdio>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include "ph_anim.h"
#include "ph_ron.h"
#include "ph_scene.h"
static std::string slurp(const char* p) { std::ifstream f(p, std::ios::binary); std::ostringstream s; s << f.rdbuf(); return s.str(); }
static int run(const std::string& text) {
    std::string err;
    ph::RonPtr root = ph::ron_parse(text, err);
    if (!root) return 1;
    ph::Scene sc;
    if (!sc.load(*root)) return 2;
    std::vector<ph::TableEntry> t;
    sc.uniform_table(t);
    ph::Player pl(sc);
    for (size_t a = 0; a < sc.animations.size() && a < 4; a++) {
        pl.init_animation_by_name(sc.animations[a].name);
        for (double tm : {0.0, 0.7, 3.3}) { pl.update(tm); sc.uniform_table(t); }
    }
    for (auto& kv : sc.stages) { pl.init_stage_by_name(kv.first); pl.update(0.2); }
    return 0;
}
int main(int argc, char** argv) {
    int ok = 0, rej = 0;
    std::mt19937 rng(12345);
    for (int i = 1; i < argc; i++) {
        const std::string base = slurp(argv[i]);
        if (run(base) == 0) ok++; else rej++;
        for (int m = 0; m < 30; m++) {   // mutations
            std::string t = base;
            int n = 1 + int(rng() % 5);
            for (int k = 0; k < n && !t.empty(); k++) {
                size_t pos = rng() % t.size();
                switch (rng() % 4) {
                    case 0: t.erase(pos, 1 + rng() % 30); break;
                    case 1: t.insert(pos, std::string(1, "()[]{},:\"'-"[rng() % 11])); break;
                    case 2: t[pos] = "()[]{},:\" 0123456789abcxyz._-"[rng() % 30]; break;
                    default: t.resize(pos); break;
                }
            }
            if (run(t) == 0) ok++; else rej++;
        }
    }
    std::printf("ok %d rejected %d\n", ok, rej);
    return 0;
}
