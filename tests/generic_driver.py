"""TEST INFRASTRUCTURE: a C++ driver generated from a machine description — fills every static array through the generated
class's accessors (whole memory box, margins included), calls the given kernels, prints every scalar and a checksum of
every array.  Used to run arbitrary OM programs through the generated host class under the sanitizers
(tests/test_race_detection.py) without writing a driver per program."""
from typing import List

CPP_TYPE = {"Int": "int", "Double": "double", "Float": "float", "Bool": "bool", "Integer": "long long int"}


def driver_source(desc: dict, kernels: List[str]) -> str:
    name, dim = desc["name"], desc["dim"]
    idx = [f"i{k}" for k in range(dim)]
    out = ["#include <cstdio>", f'#include "{name}.hpp"', "int main() {", f"  {name} sim;"]
    loops = "".join(f"  for (int i{k} = -sim.om_lower_margin_{k}(); i{k} < sim.om_size_{k}() + sim.om_upper_margin_{k}(); ++i{k})\n"
                    for k in reversed(range(dim)))
    mix = " + ".join(f"{c} * (i{k} + 7)" for k, c in zip(range(dim), (31, 17, 13)))
    arrays = [s for s in desc["statics"] if s["realm"] == "Array"]
    scalars = [s for s in desc["statics"] if s["realm"] == "Scalar"]
    for s in arrays:
        t = CPP_TYPE[s["type"]]
        val = f"({t})((({mix}) % 97) - 40)" if s["type"] in ("Int", "Integer") else f"({t})(0.01 * ((({mix}) % 97) + 1))"
        out.append(loops + f"    sim.{s['name']}({', '.join(idx)}) = {val};")
    for k in kernels:
        out.append(f"  sim.{k}();")
    for s in scalars:
        fmt = "%d" if s["type"] == "Int" else "%.17g"
        out.append(f'  printf("{s["name"]} {fmt}\\n", sim.{s["name"]}());')
    for s in arrays:
        out.append("  { double acc = 0; unsigned long long h = 1469598103934665603ull;")
        out.append(loops + f"    {{ const double v = (double)sim.{s['name']}({', '.join(idx)}); acc += v; "
                           "h = (h ^ (unsigned long long)(long long)(v * 4096.0)) * 1099511628211ull; }")
        out.append(f'    printf("{s["name"]} %.17g %llu\\n", acc, h); }}')
    out += ["  return 0;", "}"]
    return "\n".join(out) + "\n"
