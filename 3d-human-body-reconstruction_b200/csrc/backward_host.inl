// Host orchestration of the backward pass (included by smplk_api.cu).

struct BwdLayout {
  int m_blocks, n_blocks, k_blocks, splits, kbps, mpad;
  bool f16;           // backward GEMM on fp16 two-term-split operands (else 3xTF32)
  size_t off_dverts, off_dvp_hi, off_dvp_lo, off_dA, off_dtr, off_scale, off_dfeat, off_dAp, total;
};

// fp16 operands need the grouped skinning-backward kernel (it writes the scaled half rows)
static bool bwd_uses_f16(const smplk_model* mdl) {
  const ModelDev& d = mdl->d;
  return mdl->bwd_f16 && !d.lbs_only && d.grp_ok && !mdl->force_skin_v1 && ((d.V * 3) % 2 == 0);
}

static BwdLayout bwd_layout(const smplk_model* mdl, int batch) {
  const ModelDev& d = mdl->d;
  BwdLayout L;
  memset(&L, 0, sizeof(L));
  const bool pair = mdl->use_2cta && batch > kBlendBM;      // CTA-pair GEMM: 256-row m blocks
  const int bm = pair ? 2 * kBlendBM : kBlendBM;
  L.m_blocks = (batch + bm - 1) / bm;
  L.mpad = L.m_blocks * bm;
  size_t off = 0;
  L.off_dverts = off; off += align_up((size_t)batch * d.V * 3 * sizeof(float), 1024);
  L.off_dA = off;     off += align_up((size_t)batch * d.J * 12 * sizeof(float), 1024);
  L.off_dtr = off;    off += align_up((size_t)batch * 3 * sizeof(float), 1024);
  L.off_scale = off;  off += align_up((size_t)batch * 2 * sizeof(float), 1024);
  L.off_dAp = off;    off += align_up((size_t)batch * d.seg_count * 12 * sizeof(float), 1024);   // dA_seg_kernel partials
  L.f16 = bwd_uses_f16(mdl);
  if (!d.lbs_only) {
    L.n_blocks = (d.Kpad + kBlendBN - 1) / kBlendBN;
    L.k_blocks = d.Npad / ((pair ? 128 : kRowBytes) / (L.f16 ? 2 : 4));
    int splits = std::max(1, (pair ? mdl->num_sms / 2 : mdl->num_sms) / (L.m_blocks * L.n_blocks));
    splits = std::min(splits, L.k_blocks);
    L.kbps = (L.k_blocks + splits - 1) / splits;
    L.splits = (L.k_blocks + L.kbps - 1) / L.kbps;
    L.off_dvp_hi = off; off += align_up((size_t)batch * d.Npad * sizeof(float), 1024);
    L.off_dvp_lo = off; off += align_up((size_t)batch * d.Npad * sizeof(float), 1024);
    L.off_dfeat = off;  off += align_up((size_t)L.splits * L.mpad * d.Kpad * sizeof(float), 1024);
  }
  L.total = off;
  return L;
}

extern "C" size_t smplk_backward_scratch_bytes(const smplk_model* model, int32_t batch) {
  if (!model || batch < 1) return 0;
  return bwd_layout(model, batch).total;
}

extern "C" int smplk_backward(const smplk_model* model, const smplk_backward_args* a) {
  if (!model || !a) return fail(SMPLK_E_ARG, "null argument");
  const ModelDev& d = model->d;
  if (a->batch < 1 || !a->pose) return fail(SMPLK_E_ARG, "batch and pose are required");
  if (!(a->flags & (SMPLK_FLAG_SAVE_FOR_BACKWARD | SMPLK_FLAG_FIT_VERTEX_L2)))
    return fail(SMPLK_E_ARG, "backward needs the workspace of a forward run with SMPLK_FLAG_SAVE_FOR_BACKWARD");
  // after smplk_fit_vertex_l2's fused kernel d_v_posed already sits in the workspace (bf16 split rows)
  const bool dvp_ready = (a->flags & SMPLK_FLAG_FIT_VERTEX_L2) && fit_fused_applies(model);
  if (dvp_ready && (a->d_joints || a->d_joints_regressed))
    return fail(SMPLK_E_ARG, "smplk_fit_vertex_l2's backward takes d_verts only");
  if (a->betas && a->betas_batch != 1 && a->betas_batch != a->batch)
    return fail(SMPLK_E_SHAPE, "betas_batch must be 1 or batch");
  const WsLayout w = ws_layout(d, a->batch, a->flags);
  if (!a->workspace || a->workspace_bytes < w.total)
    return fail(SMPLK_E_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.total, a->workspace_bytes);
  const BwdLayout L = bwd_layout(model, a->batch);
  if (!a->scratch || a->scratch_bytes < L.total)
    return fail(SMPLK_E_WORKSPACE, "scratch too small: need %zu bytes, got %zu", L.total, a->scratch_bytes);
  if ((reinterpret_cast<uintptr_t>(a->scratch) & 255) || (reinterpret_cast<uintptr_t>(a->workspace) & 255))
    return fail(SMPLK_E_WORKSPACE, "workspace and scratch must be 256-byte aligned");
  DEVICE_GUARD(model->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(a->stream);
  const int B = a->batch;
  // Programmatic dependent launch inside this call: the FIRST kernel is launched with plain stream ordering (everything
  // the caller queued before is complete when it starts), the later ones of the fitting-step chain (dA_seg ->
  // backward GEMM -> split reduction -> pose backward) with the attribute.  The pose backward kernel recomputes its
  // forward half from the caller's read-only inputs before its wait, which that first plain launch makes safe.
  bool chain_started = false;
  auto pdl_next = [&]() { const bool p = model->use_pdl && chain_started; chain_started = true; return p; };
  if (a->d_betas && model->d.NB > 0 && (!a->betas || a->betas_batch == 1))    // shared betas: atomic accumulation target
    CUDA_TRY(cudaMemsetAsync(a->d_betas, 0, (size_t)model->d.NB * sizeof(float), st));
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  const float* A = reinterpret_cast<const float*>(ws + w.off_A);
  const float* v_posed = reinterpret_cast<const float*>(ws + w.off_vposed);
  uint8_t* sc = reinterpret_cast<uint8_t*>(a->scratch);
  float* dverts_eff = reinterpret_cast<float*>(sc + L.off_dverts);
  float* dA = reinterpret_cast<float*>(sc + L.off_dA);
  float* dtr = reinterpret_cast<float*>(sc + L.off_dtr);
  float* row_scale = reinterpret_cast<float*>(sc + L.off_scale);
  float* row_scale_inv = row_scale + B;
  float* dvp_hi = reinterpret_cast<float*>(sc + L.off_dvp_hi);
  float* dvp_lo = reinterpret_cast<float*>(sc + L.off_dvp_lo);
  float* dfeat = reinterpret_cast<float*>(sc + L.off_dfeat);
  const int joints_ld = 3 * (d.J + d.E);

  // ---- keypoint fitting: only vertex-pick joints carry a vertex gradient -> one sparse kernel
  const bool want_blend_grads = !d.lbs_only && (a->d_pose || a->d_betas || a->d_hand_pca_l || a->d_hand_pca_r);
  const bool picks_only = model->sparse_picks && !a->d_verts && !(a->d_joints_regressed && d.R > 0) && a->d_joints &&
                          d.E > 0 && !dvp_ready && (d.lbs_only || d.pick_pd != nullptr);
  if (picks_only) {
    PickBwdArgs pk;
    pk.B = B; pk.d_joints = a->d_joints; pk.joints_ld = joints_ld; pk.A = A;
    pk.vsrc = d.lbs_only ? d.bias : v_posed; pk.vsrc_stride = d.lbs_only ? 0 : (size_t)d.Npad;
    pk.dA = dA; pk.dtr = dtr; pk.d_feat = want_blend_grads ? dfeat : nullptr;
    { ProfScope prof(model, st, SMPLK_PROF_DA);
    pick_backward_kernel<<<B, kPickThreads, pick_bwd_smem_bytes(d.J, d.E), st>>>(d, pk); }
    LAUNCH_CHECK("pick_backward_kernel");
    chain_started = true;
  }

  // ---- effective vertex gradient (vertex picks and posed-vertex regressors fold into it)
  const bool scatter = !picks_only && ((a->d_joints && d.E > 0) || (a->d_joints_regressed && d.R > 0));
  const float* dverts = a->d_verts;
  if (scatter) {
    const size_t bytes = (size_t)B * d.V * 3 * sizeof(float);
    if (a->d_verts) CUDA_TRY(cudaMemcpyAsync(dverts_eff, a->d_verts, bytes, cudaMemcpyDeviceToDevice, st));
    else CUDA_TRY(cudaMemsetAsync(dverts_eff, 0, bytes, st));
    dim3 grid((d.sc_T + 127) / 128, B);
    scatter_joint_grads_kernel<<<grid, 128, 0, st>>>(d, B, (d.E > 0) ? a->d_joints : nullptr, joints_ld,
                                                     (d.R > 0) ? a->d_joints_regressed : nullptr,
                                                     dverts_eff);
    LAUNCH_CHECK("scatter_joint_grads_kernel");
    chain_started = true;
    dverts = dverts_eff;
  }

  const bool have_dv = dverts != nullptr;
  const float* seg_partials = nullptr;     // set when dA_seg_kernel ran: the pose kernel sums its per-segment partials
  if (picks_only) {
    // dA, dtr and d_feat are already there
  } else if (have_dv && dvp_ready && d.w_rows_normalised && d.seg_count > 0 && !model->da_v1) {
    // fitting step: segment kernel + fixed-order reduction (d_transl from the translation columns)
    DASegArgs ds;
    ds.B = B; ds.dverts = dverts; ds.vsrc = v_posed; ds.vsrc_stride = (size_t)d.Npad;
    ds.dAp = reinterpret_cast<float*>(sc + L.off_dAp);
    ds.bodies_per_warp = std::max(1, std::min(16, B / 8));
    { ProfScope prof(model, st, SMPLK_PROF_DA);
    dim3 grid((d.seg_count + kDASegWarps - 1) / kDASegWarps, (B + ds.bodies_per_warp - 1) / ds.bodies_per_warp);
    // paired gathers need 8-byte aligned rows: (B, V, 3) gradient rows (V even + aligned base) and v_posed rows (Npad % 4 == 0)
    const bool pair = (d.V % 2 == 0) && (reinterpret_cast<uintptr_t>(dverts) % 8 == 0) &&
                      (reinterpret_cast<uintptr_t>(v_posed) % 8 == 0) && (d.Npad % 2 == 0);
    const bool pdl = pdl_next();
    if (pair) launch_k(pdl, dA_seg_kernel<true>, grid, kDASegWarps * 32, 0, st, d, ds);
    else launch_k(pdl, dA_seg_kernel<false>, grid, kDASegWarps * 32, 0, st, d, ds); }
    LAUNCH_CHECK("dA_seg_kernel");
    seg_partials = ds.dAp;
  } else if (have_dv) {
    DAArgs da;
    da.B = B; da.dverts = dverts;
    da.vsrc = d.lbs_only ? d.bias : v_posed;
    da.vsrc_stride = d.lbs_only ? 0 : (size_t)d.Npad;
    da.dA = dA; da.dtr = dtr;
    da.row_scale = (L.f16 && !dvp_ready) ? row_scale : nullptr; da.row_scale_inv = (L.f16 && !dvp_ready) ? row_scale_inv : nullptr;
    { ProfScope prof(model, st, SMPLK_PROF_DA);
    int jsplit = 1;
    while (jsplit < 8 && (long)B * jsplit * 2 <= 4L * model->num_sms && jsplit * (kDAThreads / 32) < d.J) jsplit *= 2;
    dA_kernel<<<dim3(B, jsplit), kDAThreads, 0, st>>>(d, da); }
    LAUNCH_CHECK("dA_kernel");
    chain_started = true;
  } else {
    CUDA_TRY(cudaMemsetAsync(dA, 0, (size_t)B * d.J * 12 * sizeof(float), st));
    CUDA_TRY(cudaMemsetAsync(dtr, 0, (size_t)B * 3 * sizeof(float), st));
  }

  const bool blend_bwd = have_dv && !d.lbs_only && (a->d_pose || a->d_betas || a->d_hand_pca_l || a->d_hand_pca_r);
  if (blend_bwd) {
    if (!model->has_tma) return fail(SMPLK_E_DEVICE, "tcgen05 path unavailable");
    SkinBwdArgs sb;
    sb.B = B; sb.dverts = dverts; sb.A = A; sb.dvp_hi = dvp_hi; sb.dvp_lo = dvp_lo;
    sb.h_hi = reinterpret_cast<__half*>(dvp_hi); sb.h_lo = reinterpret_cast<__half*>(dvp_lo);
    sb.row_scale = row_scale;
    const bool f16 = L.f16;
    const int tiles = (d.V + kSkinTileVerts - 1) / kSkinTileVerts;
    const bool grouped = d.grp_ok && !model->force_skin_v1 && ((d.V * 3) % 2 == 0);
    if (dvp_ready) {
      dvp_hi = reinterpret_cast<float*>(ws + w.off_dvp);
      dvp_lo = reinterpret_cast<float*>(ws + w.off_dvp + (size_t)w.chunk * d.Npad * sizeof(uint16_t));
    } else if (grouped) {
      constexpr int kS = kSkinBwdStages;
      int bpb = pick_bpb(B, tiles, 2 * model->num_sms, 32);
      if (model->skin_bpb > 0) bpb = model->skin_bpb;
      sb.bodies_per_block = bpb;
      dim3 grid(tiles, (B + bpb - 1) / bpb);
      const size_t smem = (size_t)((kS + 1) * kSkinTileVerts * 3 + 2 * kGrpABodies * grp_a_pad(d.J)) * sizeof(float);
      ProfScope prof(model, st, SMPLK_PROF_SKIN_BWD);
      if (f16) skin_backward_grouped_kernel<kS, true><<<grid, kGrpThreads, smem, st>>>(d, sb);
      else skin_backward_grouped_kernel<kS, false><<<grid, kGrpThreads, smem, st>>>(d, sb);
      LAUNCH_CHECK("skin_backward_grouped_kernel");
      chain_started = true;
    } else {
      int bpb = 16;
      while (bpb > 1 && (long)tiles * ((B + bpb - 1) / bpb) < 4L * 4 * model->num_sms) bpb >>= 1;
      sb.bodies_per_block = bpb;
      dim3 grid(tiles, (B + bpb - 1) / bpb);
      const size_t smem = (size_t)(3 * kSkinTileVerts * 3 + d.J * 12) * sizeof(float);
      { ProfScope prof(model, st, SMPLK_PROF_SKIN_BWD);
      if (d.ell_k <= 4) skin_backward_kernel<true><<<grid, kSkinThreads, smem, st>>>(d, sb);
      else skin_backward_kernel<false><<<grid, kSkinThreads, smem, st>>>(d, sb); }
      LAUNCH_CHECK("skin_backward_kernel");
      chain_started = true;
    }

    CUtensorMap tm_ahi, tm_alo, tm_out;
    const bool pair = model->use_2cta && B > kBlendBM;
    if (pair) {
      if (int r = make_operand_tmap_2cta(model, &tm_ahi, dvp_hi, d.Npad, B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
      if (int r = make_operand_tmap_2cta(model, &tm_alo, dvp_lo, d.Npad, B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
      if (int r = make_tmap_2d(model, &tm_out, dfeat, d.Kpad, (uint64_t)L.splits * L.mpad, kEpiCols, kBlendBM,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE)) return r;
      BlendGemmArgs ga;
      ga.num_m_blocks = L.m_blocks; ga.num_n_blocks = L.n_blocks; ga.num_k_blocks = L.k_blocks;
      ga.num_splits = L.splits; ga.k_blocks_per_split = L.kbps; ga.out_rows_per_split = L.mpad;
      ga.k_elems = d.Npad; ga.out_scale = f16 ? 1.0f / d.pd_scale : 1.0f;
      ga.bias = nullptr;
      ga.row_scale = (f16 && !dvp_ready) ? row_scale_inv : nullptr; ga.row_scale_rows = B;
      ga.a_bf16 = dvp_ready ? 1 : 0;
      ga.out = dfeat; ga.out_ld = d.Kpad; ga.out_rows = L.splits * L.mpad; ga.out_cols = d.Kpad;
      const int tiles_g = L.m_blocks * L.n_blocks * L.splits;
      ProfScope prof(model, st, SMPLK_PROF_BLEND_BWD);
      const bool pdl = pdl_next();
      if (f16)
        launch_k(pdl, blend_tcgen05_2cta_kernel<true>, 2 * std::min(tiles_g, model->num_sms / 2), kGemmThreads, k2SmemAlloc, st,
                 tm_ahi, tm_alo, dvp_ready ? model->tmap2_pdknb_hi : model->tmap2_pdknh_hi,
                 dvp_ready ? model->tmap2_pdknb_lo : model->tmap2_pdknh_lo, tm_out, ga);
      else
        launch_k(pdl, blend_tcgen05_2cta_kernel<false>, 2 * std::min(tiles_g, model->num_sms / 2), kGemmThreads, k2SmemAlloc, st,
                 tm_ahi, tm_alo, model->tmap2_pdkn_hi, model->tmap2_pdkn_lo, tm_out, ga);
      LAUNCH_CHECK("blend_tcgen05_2cta_kernel(backward)");
    } else {
    if (int r = make_operand_tmap(model, &tm_ahi, dvp_hi, d.Npad, B, kBlendBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
    if (int r = make_operand_tmap(model, &tm_alo, dvp_lo, d.Npad, B, kBlendBM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, f16)) return r;
    if (int r = make_tmap_2d(model, &tm_out, dfeat, d.Kpad, (uint64_t)L.splits * L.mpad, kEpiCols, kBlendBM,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE)) return r;
    BlendGemmArgs ga;
    ga.num_m_blocks = L.m_blocks; ga.num_n_blocks = L.n_blocks; ga.num_k_blocks = L.k_blocks;
    ga.num_splits = L.splits; ga.k_blocks_per_split = L.kbps; ga.out_rows_per_split = L.mpad;
    ga.k_elems = d.Npad; ga.out_scale = f16 ? 1.0f / d.pd_scale : 1.0f;
    ga.bias = nullptr;
    ga.row_scale = (f16 && !dvp_ready) ? row_scale_inv : nullptr; ga.row_scale_rows = B;
    ga.a_bf16 = dvp_ready ? 1 : 0;
    ga.out = dfeat; ga.out_ld = d.Kpad; ga.out_rows = L.splits * L.mpad; ga.out_cols = d.Kpad;
    const int tiles_g = L.m_blocks * L.n_blocks * L.splits;
    { ProfScope prof(model, st, SMPLK_PROF_BLEND_BWD);
    if (f16)
      blend_tcgen05_kernel<true><<<std::min(tiles_g, model->num_sms), kGemmThreads, kGemmSmemAlloc, st>>>(
          tm_ahi, tm_alo, dvp_ready ? model->tmap_pdknb_hi : model->tmap_pdknh_hi,
          dvp_ready ? model->tmap_pdknb_lo : model->tmap_pdknh_lo, tm_out, ga);
    else
      blend_tcgen05_kernel<false><<<std::min(tiles_g, model->num_sms), kGemmThreads, kGemmSmemAlloc, st>>>(
          tm_ahi, tm_alo, model->tmap_pdkn_hi, model->tmap_pdkn_lo, tm_out, ga); }
    LAUNCH_CHECK("blend_tcgen05_kernel(backward)");
    chain_started = true;
    }
  }

  PoseBwdArgs pb;
  pb.B = B; pb.betas = a->betas; pb.betas_B = a->betas ? a->betas_batch : 1;
  pb.pose = a->pose; pb.pca_l = a->hand_pca_l; pb.pca_r = a->hand_pca_r;
  pb.add_mean = (a->flags & SMPLK_FLAG_ADD_POSE_MEAN) ? 1 : 0;
  pb.dA = dA; pb.d_joints = a->d_joints; pb.joints_ld = joints_ld;
  pb.d_feat = (blend_bwd || (picks_only && want_blend_grads)) ? dfeat : nullptr;
  pb.feat_splits = picks_only ? 1 : L.splits; pb.feat_split_stride = (size_t)L.mpad * d.Kpad;
  if (pb.d_feat != nullptr && L.splits > 1 && !picks_only) {      // sum the split-K partials in parallel, in place
    const int n = B * d.Kpad;
    launch_k(pdl_next(), reduce_splits_kernel, (n + 255) / 256, 256, 0, st, n, L.splits, pb.feat_split_stride, dfeat);
    LAUNCH_CHECK("reduce_splits_kernel");
    pb.feat_splits = 1;
  }
  pb.dtr_verts = seg_partials ? nullptr : dtr;
  pb.dAp = seg_partials;
  pb.d_betas = d.NB > 0 ? a->d_betas : nullptr;
  pb.d_pose = a->d_pose; pb.d_pca_l = a->d_hand_pca_l; pb.d_pca_r = a->d_hand_pca_r;
  pb.d_transl = a->d_transl;
  pb.d_loss = a->d_loss;
  pb.d_loss_stride = (a->flags & SMPLK_FLAG_LOSS_SUM) ? 0 : 1;
  pb.d_full_pose = a->d_full_pose;
  // (a shared-betas gradient was zeroed at the top of the call: no memset node between the chain's kernels)
  const int blocks = (B + kPoseWarps - 1) / kPoseWarps;
  pb.staged_segs = 0;
  if (seg_partials && (size_t)kPoseWarps * pose_bwd_smem_floats(d.J, d.Kpad, d.seg_count) * sizeof(float) <= 96 * 1024)
    pb.staged_segs = d.seg_count;           // two blocks per SM still fit
  const size_t smem = (size_t)kPoseWarps * pose_bwd_smem_floats(d.J, d.Kpad, pb.staged_segs) * sizeof(float);
  { ProfScope prof(model, st, SMPLK_PROF_POSE_BWD);
  const bool pdl = pdl_next();
  if (d.J <= 32) launch_k(pdl, pose_backward_kernel<1>, blocks, kPoseWarps * 32, smem, st, d, pb);
  else launch_k(pdl, pose_backward_kernel<2>, blocks, kPoseWarps * 32, smem, st, d, pb); }
  LAUNCH_CHECK("pose_backward_kernel");
  return 0;
}
