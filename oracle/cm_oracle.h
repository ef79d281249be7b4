/* cm_oracle.h -- C interface of the CPU oracle (test infrastructure only). */
#ifndef CM_ORACLE_H
#define CM_ORACLE_H

#define ORC_MAX_P 64
#define ORC_MAX_SLOTS 1024

typedef struct {
    int B, H, W, P, F;   /* batch, resolution, passes_loss, number of flow scales   */
    int S;               /* scales_loss                                             */
    int mode;            /* iterative_mode: 1 = one, 2 = two, 4 = four              */
    int border_comp;     /* BaseEventWarping.border_compensation                    */
    int loss_scaling;    /* BaseEventWarping.loss_scaling                           */
    int round_ts;        /* config["loss"]["round_ts"]                              */
} orc_cfg;

int orc_num_slots(const orc_cfg *c, int linear);
int orc_version(void);
int orc_set_threads(int n);   /* returns the OpenMP team size in effect */

#endif
