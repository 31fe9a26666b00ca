/* placeholder: replaced by the full standard model below */
#include "orc_node.h"
int orc_std_active_list(const orc_evolve_ctx *c, int *active) { (void)c; (void)active; return 0; }
void orc_std_scales(orc_evolve_ctx *c, double *s) { (void)c; (void)s; }
void orc_std_solve_analytics(orc_evolve_ctx *c, double time) { (void)c; (void)time; }
int orc_std_rates(orc_evolve_ctx *c, double time, double *rate) { (void)c; (void)time; (void)rate; return 0; }
void orc_std_post_step(orc_evolve_ctx *c, int *status) { (void)c; (void)status; }
void orc_std_post_evolve(orc_evolve_ctx *c) { (void)c; }
void orc_std_pre_evolve(orc_evolve_ctx *c) { (void)c; }
