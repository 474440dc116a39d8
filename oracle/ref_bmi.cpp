// ref_bmi.cpp — TEST INFRASTRUCTURE (oracle/): the REFERENCE's bit-interleaving primitives run here.
// shammath/include/shammath/sfc/bmi.hpp (expand_bits / contract_bits, the core of its Morton and Hilbert codes) is
// plain C++; this driver is compiled against it where it lies (oracle/Makefile, target _ref/bmi_ref).
//   bmi_ref < numbers     one line per input x: expand_bits<u32,2>(x) expand_bits<u64,2>(x) contract_bits<u32,2>(x)
//                         contract_bits<u64,2>(x)
#include "shammath/sfc/bmi.hpp"
#include <cstdio>
#include <iostream>

int main() {
    using namespace shamrock::sfc::bmi;
    unsigned long long x;
    while (std::cin >> x)
        std::printf("%u %llu %u %llu\n", expand_bits<u32, 2>(u32(x)), (unsigned long long) expand_bits<u64, 2>(u64(x)),
                    contract_bits<u32, 2>(u32(x)), (unsigned long long) contract_bits<u64, 2>(u64(x)));
    return 0;
}
