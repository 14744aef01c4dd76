// compile probe: the reference's own solver templates on top of the B200 header layer
#include "hala.hpp"
#include "hala_solvers.hpp"
#include <iostream>
int main(){
    hala::gpu_engine engine(0);
    std::vector<int> pntr = {0, 2, 5, 8, 11, 13}, indx = {0, 1, 0, 1, 2, 1, 2, 3, 2, 3, 4, 3, 4};
    std::vector<double> vals = {2.0, 1.0, 1.0, 2.0, 1.0, 1.0, 2.0, 1.0, 1.0, 2.0, 1.0, 1.0, 2.0}, b = {4, 8, 12, 16, 14};
    auto gp = engine.load(pntr); auto gi = engine.load(indx); auto gv = engine.load(vals); auto gb = engine.load(b);
    hala::gpu_vector<double> gx(engine.device());
    int it = hala::solve_cg(engine, hala::stop_criteria<double>(1.E-9, 100), gp, gi, gv,
                            [&](auto const &in, auto &out)->void{ hala::vcopy(engine, in, out); }, gb, gx);
    auto x = gx.unload();
    std::cout << it << " " << x[0] << " " << x[4] << std::endl;
    return 0;
}
