/* kgo_env.c -- env layer of the fp64 oracle (test infrastructure): KinovaGripper_Env.step / _get_obs /
 * _get_reward / reset (kinova_gripper_env.py:1495-1552, 438-534, 631-687, 1310-1410) and the expert
 * controller + its episode bookkeeping (expert_data.py:318-352, 367-671, 745-804). */
#include "kgo_internal.h"

typedef struct { double R[9], t[3], wrist[3]; } HandFrame;

/* _get_trans_mat_wrist_pose, ENV:274-288 */
static void hand_frame(const KgoData *d, HandFrame *h) {
    static const double C[9] = {0, 0, 1, -1, 0, 0, 0, -1, 0};
    double tmp[9];
    kgo_mulmm3(tmp, d->geom_xmat + 9, C);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) h->R[3 * i + j] = tmp[3 * j + i];
    double off[3] = {-0.009, 0.048, 0.0}, w[3];
    multmv3(w, h->R, off);
    add3(h->wrist, d->geom_xpos + 3, w);
    mulmv3(h->t, h->R, h->wrist);
    scl3(h->t, h->t, -1);
}
static void to_local(const HandFrame *h, const double *p, double *out) {
    double r[3];
    mulmv3(r, h->R, p);
    add3(out, r, h->t);
}

/* _get_dot_product, ENV:591-608 */
static double dot_product(const KgoData *d, const double *p) {
    const double *hand = d->xpos + 3;
    double ox = fabs(p[0] - hand[0]), oy = fabs(p[1] - hand[1]);
    double on = sqrt(ox * ox + oy * oy);
    double cx = fabs(0.0 - hand[0]), cy = fabs(0.0 - hand[1]);
    double cn = sqrt(cx * cx + cy * cy);
    double dp = (ox / on) * (cx / cn) + (oy / on) * (cy / cn);
    return pow(dp, 20);
}

/* _get_obj_size, ENV:706-746 (single-geom objects) */
static void obj_size(const KgoModel *m, int shape, double *out) {
    double size[3];
    copy3(size, m->obj_size[shape]);
    if (size[2] == 0) { size[2] = size[1]; size[1] = size[0]; }
    double diffs[3] = {fabs(size[0] - size[1]), fabs(size[1] - size[2]), fabs(size[0] - size[2])};
    int am = 0;
    if (diffs[1] < diffs[am]) am = 1;
    if (diffs[2] < diffs[am]) am = 2;
    if (am != 0 || (m->obj_flags[shape] & 1)) { double t = size[0]; size[0] = size[2]; size[2] = t; }
    out[0] = size[0] > 0 ? size[0] : 0; out[1] = size[1] > 0 ? size[1] : 0; out[2] = size[2];
}

static double tri_area(const double *a, const double *b) {
    double c[3];
    cross3(c, a, b);
    return 0.5 * norm3(c);
}

void kgo_get_obs(const KgoModel *m, const KgoState *s, const KgoData *d, int global_rep, double *obs) {
    static const int fgeom[6] = {2, 4, 6, 3, 5, 7}; /* f1_prox f2_prox f3_prox f1_dist f2_dist f3_dist */
    HandFrame h;
    hand_frame(d, &h);
    const double *obj = d->geom_xpos + 24;
    double loc_obj[3], fdp[6], joint[9], size[3], fod[12], range[17];
    to_local(&h, obj, loc_obj);
    /* _get_angles, ENV:569-582 */
    double ow[3]; copy3(ow, loc_obj); { double n = norm3(ow); ow[0] /= n; ow[1] /= n; ow[2] /= n; }
    double z_angle = acos(ow[1] / sqrt(ow[0] * ow[0] + ow[1] * ow[1]));
    double x_angle = acos(ow[1] / sqrt(ow[1] * ow[1] + ow[2] * ow[2]));
    for (int i = 0; i < 9; i++) joint[i] = d->sensordata[i];
    joint[0] = -joint[0]; joint[1] = -joint[1];
    obj_size(m, s->shape, size);
    for (int i = 0; i < 12; i++) {
        double df[3];
        sub3(df, d->site_xpos + 3 * m->fod_site[i], obj);
        fod[i] = norm3(df);
    }
    for (int i = 0; i < 17; i++) { double r = d->sensordata[9 + i]; range[i] = (r == -1) ? 6 : r; }
    for (int i = 0; i < 6; i++) fdp[i] = dot_product(d, d->geom_xpos + 3 * fgeom[i]);
    double dp = dot_product(d, obj);
    int n = 0;
    if (global_rep) {
        for (int i = 0; i < 6; i++) for (int c = 0; c < 3; c++) obs[n++] = d->geom_xpos[3 * fgeom[i] + c];
        for (int c = 0; c < 3; c++) obs[n++] = h.wrist[c];
        for (int c = 0; c < 3; c++) obs[n++] = obj[c];
        for (int i = 0; i < 9; i++) obs[n++] = joint[i];
        obs[n++] = size[0]; obs[n++] = size[1]; obs[n++] = size[2] * 2;
        for (int i = 0; i < 12; i++) obs[n++] = fod[i];
        obs[n++] = x_angle; obs[n++] = z_angle;
        for (int i = 0; i < 17; i++) obs[n++] = range[i];
        for (int i = 0; i < 6; i++) obs[n++] = fdp[i];
        obs[n++] = dp;
        return; /* 74 */
    }
    double fl[18], wl[3], grav[3], gw[3] = {0, 0, -1};
    for (int i = 0; i < 6; i++) to_local(&h, d->geom_xpos + 3 * fgeom[i], fl + 3 * i);
    to_local(&h, h.wrist, wl);
    mulmv3(grav, h.R, gw);
    /* experimental_sensor, ENV:290-343 */
    double s1[3], s2[3];
    sub3(s1, fl, fl + 6); sub3(s2, fl, fl + 3);
    double front_area = tri_area(s1, s2);
    double top1 = tri_area(fl, fl + 9), top2 = tri_area(fl + 9, fl + 12), top3 = tri_area(fl + 3, fl + 12);
    double top4 = tri_area(fl + 6, fl + 15), top5 = tri_area(fl + 9, fl + 15);
    double total1 = top1 + top2 + top3, total2 = top1 + top4 + top5;
    double top_area = total1 > total2 ? total1 : total2;
    double sp[3] = {0, 0, 0}; int ns = 0;
    for (int i = 0; i < 5; i++) {
        double t[3];
        to_local(&h, d->site_xpos + 3 * i, t);
        if (range[i] < 0.06) { t[1] += range[i]; add3(sp, sp, t); ns++; }
    }
    if (ns == 0) { sp[0] = sp[1] = sp[2] = 0.2; } else { scl3(sp, sp, 1.0 / ns); }
    int gi = 0;
    if (fabs(grav[1]) > fabs(grav[gi])) gi = 1;
    if (fabs(grav[2]) > fabs(grav[gi])) gi = 2;
    double front_part, top_part;
    if (gi == 2) { front_part = fabs(size[0] * size[2]) / front_area; top_part = fabs(size[0] * size[1]) / top_area; }
    else if (gi == 1) { front_part = fabs(size[0] * size[2]) / front_area; top_part = fabs(size[1] * size[2]) / top_area; }
    else { front_part = fabs(size[0] * size[1]) / front_area; top_part = fabs(size[0] * size[2]) / top_area; }
    for (int i = 0; i < 18; i++) obs[n++] = fl[i];
    for (int c = 0; c < 3; c++) obs[n++] = wl[c];
    for (int c = 0; c < 3; c++) obs[n++] = loc_obj[c];
    for (int i = 0; i < 9; i++) obs[n++] = joint[i];
    obs[n++] = size[0]; obs[n++] = size[1]; obs[n++] = size[2] * 2;
    for (int i = 0; i < 12; i++) obs[n++] = fod[i];
    obs[n++] = x_angle; obs[n++] = z_angle;
    for (int i = 0; i < 17; i++) obs[n++] = range[i];
    for (int c = 0; c < 3; c++) obs[n++] = grav[c];
    for (int c = 0; c < 3; c++) obs[n++] = sp[c];
    obs[n++] = front_part; obs[n++] = top_part;
    for (int i = 0; i < 6; i++) obs[n++] = fdp[i];
    obs[n++] = dp; /* 82 */
}

/* ENV:1501-1534: ctrl from the 6-d action [slide_x, slide_y, slide_z, f1, f2, f3] */
void kgo_ctrl_from_action6(const KgoData *d, const double action[6], double ctrl[KGO_NU]) {
    HandFrame h;
    hand_frame(d, &h);
    double g[3] = {0, 0, 0.733 * 10 / 25}, stuff[3], sv[3];
    mulmv3(stuff, h.R, g);                 /* ENV:1511-1515 */
    stuff[0] = -stuff[0]; stuff[1] = -stuff[1];
    mulmv3(sv, h.R, action);               /* ENV:1518-1522 (both branches negate x and y) */
    sv[0] = -sv[0]; sv[1] = -sv[1];
    for (int i = 0; i < 3; i++) { ctrl[2 * i] = sv[i]; ctrl[2 * i + 1] = stuff[i]; ctrl[6 + i] = action[3 + i]; }
}
/* the 4-d action [wrist, f1, f2, f3] is padded to [0, 0, wrist, f1, f2, f3] (ENV:1502-1503) */
void kgo_ctrl_from_action(const KgoData *d, const double action[4], double ctrl[KGO_NU]) {
    double a6[6] = {0, 0, action[0], action[1], action[2], action[3]};
    kgo_ctrl_from_action6(d, a6, ctrl);
}
/* determine_hand_location (ENV:1286-1307): slide start positions from the Tfw of the PREVIOUS episode (d_prev = its last
 * forward pass; NULL = a freshly constructed env, Tfw = 0).  orient: 0 normal, 1 rotated, 2 top; obj_size 'm' (ENV:62). */
void kgo_hand_location(const KgoData *d_prev, int orient, double hand[3]) {
    hand[0] = hand[1] = hand[2] = 0;
    if (orient == 0 || !d_prev) return;
    HandFrame h;
    hand_frame(d_prev, &h);
    double vt[3] = {-0.005, -0.155, 0.14 + 0.06}, vr[3] = {0.051, -0.075, 0.06}, loc[3];
    mulmv3(loc, h.R, orient == 2 ? vt : vr);
    hand[0] = -loc[0]; hand[1] = -loc[1]; hand[2] = loc[2];
}

void kgo_env_reset(const KgoModel *m, KgoEnv *e, int shape, const double hand_euler[3], const double obj_pos[3], double obs[KGO_NOBS]) {
    kgo_env_reset_pose(m, e, shape, hand_euler, obj_pos, NULL, obs);
}
/* hand_pos = qpos[0:3] as _set_state (ENV:692-703) receives them from determine_hand_location; NULL = zeros */
void kgo_env_reset_pose(const KgoModel *m, KgoEnv *e, int shape, const double hand_euler[3], const double obj_pos[3], const double *hand_pos, double obs[KGO_NOBS]) {
    double ctrl[KGO_NU] = {0};
    memset(e, 0, sizeof *e);
    kgo_state_init(m, &e->st, shape, hand_euler, obj_pos);
    if (hand_pos) for (int c = 0; c < 3; c++) e->st.qpos[c] = hand_pos[c];
    kgo_forward(m, &e->st, ctrl, KGO_RAYS, &e->d);
    /* centroid fix-up, ENV:1382-1389 */
    double del[3];
    sub3(del, obj_pos, e->d.geom_xpos + 24);
    if (norm3(del) > 0.05) {
        for (int c = 0; c < 3; c++) e->st.qpos[9 + c] = obj_pos[c] + del[c];
        kgo_forward(m, &e->st, ctrl, KGO_RAYS, &e->d);
    }
    kgo_get_obs(m, &e->st, &e->d, 0, obs);
    e->elapsed = 0; e->max_steps = 30;
    e->init_obj_x = obs[21]; e->init_dot = obs[81];
    e->have_prev = 0; e->num_good_grasps = 0; e->total_steps = 0;
}

static int g_rays_every_substep = 0;
/* 1 = compute the 17 rangefinders in every substep, as mj_step does (only the last values are ever read, so results
 * are identical); used by the CPU-baseline timing so that the oracle carries the reference's cost structure */
void kgo_set_rays_every_substep(int on) { g_rays_every_substep = on; }

int kgo_env_step(const KgoModel *m, KgoEnv *e, const double action[4], double obs[KGO_NOBS], double *reward, double info3[3]) {
    double a6[6] = {0, 0, action[0], action[1], action[2], action[3]};
    return kgo_env_step6(m, e, a6, obs, reward, info3);
}
int kgo_env_step6(const KgoModel *m, KgoEnv *e, const double action[6], double obs[KGO_NOBS], double *reward, double info3[3]) {
    double ctrl[KGO_NU];
    kgo_ctrl_from_action6(&e->d, action, ctrl);
    for (int k = 0; k < 15; k++) kgo_step(m, &e->st, ctrl, (k == 14 || g_rays_every_substep) ? KGO_RAYS : 0, &e->d);
    kgo_get_obs(m, &e->st, &e->d, 0, obs);
    /* _get_reward, ENV:661-666 on global obs[23] = object geom z */
    double z = e->d.geom_xpos[24 + 2];
    int done = (fabs(z - 0.2) < 0.005) || (z >= 0.2);
    double lift = done ? 50.0 : 0.0;
    *reward = lift;
    if (info3) { info3[0] = 0; info3[1] = 0; info3[2] = lift; }
    e->elapsed++;
    if (e->elapsed >= e->max_steps) done = 1;
    return done;
}

/* ------------------------------------------------------------------ expert controller */
static double pid_touch_vel(double a, double b) { double err = a - b; return err + err / 15.0; }
static double pid_velocity(double dot) {
    double err = 1 - dot, vel = err + err / 15.0, act = (vel / 1.25) * 0.3;
    return act < 0.05 ? 0.05 : act;
}
static void clamp_vel(double *a) { /* check_vel_in_range, EXP:540-550 */
    for (int i = 1; i < 4; i++) { if (a[i] < 0.5) a[i] = 0.5; else if (a[i] > 0.8) a[i] = 0.8; }
}
static void pd_controller(const KgoEnv *e, const double *obs, int lift, double *a) {
    const double cv = 0.5, minv = 0.5, flv = 0.5, wlv = 0.6;
    double dot = obs[81], f1 = 0, f2 = 0, f3 = 0, wrist = 0;
    if (fabs(e->init_obj_x) <= 0.03) {
        f1 = f2 = f3 = cv;
        if (fabs(dot - e->init_dot) > 0.01) { f1 = cv; f2 = f3 = cv / 2; }
        if (lift) { wrist = wlv; f1 = flv / 2; f2 = f3 = flv; }
    } else if (e->init_obj_x > 0.0) {
        if (fabs(dot - e->init_dot) < 0.01) { f1 = 0; f2 = f3 = pid_touch_vel(dot, obs[79]); }
        else {
            if (fabs(1 - dot) > 0.01) { f1 = minv; f2 = f3 = pid_velocity(dot); }
            else { f1 = pid_touch_vel(dot, obs[78]); f2 = f3 = 0; }
            if (lift) { wrist = wlv; f1 = flv / 2; f2 = f3 = flv; }
        }
    } else {
        if (fabs(dot - e->init_dot) < 0.01) { f1 = pid_touch_vel(dot, obs[78]); f2 = f3 = 0; }
        else {
            if (fabs(1 - dot) > 0.01) { f1 = pid_velocity(dot); f2 = f3 = minv; }
            else { f2 = f3 = pid_touch_vel(dot, obs[79]); f1 = 0; }
            if (lift) { wrist = wlv; f1 = flv / 2; f2 = f3 = flv; }
        }
    }
    a[0] = wrist; a[1] = f1; a[2] = f2; a[3] = f3;
    clamp_vel(a);
}
static void naive_controller(int lift, double *a) {
    a[0] = 0; a[1] = a[2] = a[3] = 0.5;
    if (lift) { a[0] = 0.6; a[1] = a[2] = a[3] = 0.5; }
}

void kgo_expert_action(KgoEnv *e, const double obs[KGO_NOBS], int mode, double action[4], int *lift_check) {
    /* check_grasp, EXP:559-593 on obs[9], obs[12], obs[15] */
    int grasp = 0;
    if (e->have_prev) {
        double ch = fabs(e->prev_fx[0] - obs[9]) / 15 + fabs(e->prev_fx[1] - obs[12]) / 15 + fabs(e->prev_fx[2] - obs[15]) / 15;
        grasp = ch < 0.0002;
    }
    if (grasp) e->num_good_grasps++;
    int lift = (e->total_steps > 10) && (e->num_good_grasps >= 1);
    double x = obs[21];
    if (mode == 0) naive_controller(lift, action);
    else if (mode == 1) pd_controller(e, obs, lift, action);
    else {
        if (x < -0.04 || x > 0.04) pd_controller(e, obs, lift, action);
        else if ((-0.04 <= x && x <= -0.02) || (0.02 <= x && x <= 0.04)) {
            /* np.interp(arange(1,4), naive[1:3]=[.5,.5], expert[1:3]) returns expert[2] three times (EXP:654) */
            double ex[4];
            pd_controller(e, obs, lift, ex);
            action[0] = 0; action[1] = action[2] = action[3] = ex[2];
        } else naive_controller(lift, action);
    }
    action[0] = lift ? 0.6 : 0.0;
    if (lift_check) *lift_check = lift;
    /* loop bookkeeping, EXP:800-804: prev_obs only starts tracking after the first step */
    if (e->total_steps > 0) { e->prev_fx[0] = obs[9]; e->prev_fx[1] = obs[12]; e->prev_fx[2] = obs[15]; e->have_prev = 1; }
    e->total_steps++;
}
