// B200-native ASTC block compressor: partition search (astcenc_find_best_partitioning.cpp).
// k-means on lanes-over-texels with ordered chains for the centre sums, popcount mismatch on
// lanes-over-partitionings, a warp-parallel *stable* counting sort, then one lane per candidate
// partitioning for the line-error evaluation (each candidate is an independent scalar job).
#pragma once

struct PartScratch {
	SPtr<float> dist;          // [T]
	SPtr<float> cand_err;      // [L][2]
	SPtr<uint64_t> bitmaps;    // [4]
	SPtr<uint16_t> order;      // [L]
	SPtr<uint16_t> hist;       // [64]
	SPtr<uint8_t> pot;         // [T] texel -> k-means partition
	SPtr<uint8_t> mism;        // [n]
};

ASTC_FN PartScratch make_part_scratch(const WCtx& w, unsigned int L) {
	PartScratch s;
	uint32_t p = su_of(w);
	s.dist = sptr<float>(p);
	p += tp4(w);
	s.cand_err = sptr<float>(p);
	p += 8 * L;
	p = (p + 7u) & ~7u;
	s.bitmaps = sptr<uint64_t>(p);
	p += 32;
	s.order = sptr<uint16_t>(p);
	p += 2 * ((L + 1) & ~1u);
	s.hist = sptr<uint16_t>(p);
	p += 128;
	s.pot = sptr<uint8_t>(p);
	p += tp4(w) >> 2;
	s.mism = sptr<uint8_t>(p);
	return s;
}

// k-means cluster centres live in the arena: tmpf[32..48) as four f4
ASTC_FN SPtr<f4> centers_of(const WCtx& w) { return sptr<f4>(w.base + A_TMPF + 128); }

// kmeans_init :60-143
ASTC_COOP void kmeans_init(WCtx w, unsigned int partition_count, unsigned int L) {
	PartScratch ps = make_part_scratch(w, L);
	SPtr<f4> cluster_centers = centers_of(w);
	SPtr<float> tmpf = tmpf_of(w);
	int T = w.T;
	f4 cw = bi_of(w).channel_weight;
	unsigned int clusters_selected = 0;
	unsigned int sample = 145897 % (unsigned int)T;
	f4 center_color = texel4(w, (int)sample);
	if (w.lane == 0) {
		cluster_centers[0] = center_color;
	}
	clusters_selected++;
	const float cluster_cutoffs[9] = {0.626220f, 0.932770f, 0.275454f, 0.318558f, 0.240113f, 0.009190f, 0.347661f, 0.731960f, 0.156391f};
	unsigned int cutoff = (clusters_selected - 1) + 3 * (partition_count - 2);
	bool first = true;
	while (true) {
		ASTC_NOUNROLL
		for (int i = w.lane; i < T; i += ASTC_WARP) {
			f4 diff = texel4(w, i) - center_color;
			float distance = dot_s(diff * diff, cw);
			if (!first) {
				distance = minf(distance, ps.dist[i]);
			}
			ps.dist[i] = distance;
		}
		wsync();
		if (w.lane == 0) {
			float distance_sum = 0.0f;
			ASTC_NOUNROLL
			for (int i = 0; i < T; i++) {
				distance_sum += ps.dist[i];
			}
			float summa = 0.0f;
			float distance_cutoff = distance_sum * cluster_cutoffs[cutoff];
			unsigned int s;
			ASTC_NOUNROLL
			for (s = 0; s < (unsigned int)T; s++) {
				summa += ps.dist[s];
				if (summa >= distance_cutoff) {
					break;
				}
			}
			s = s < (unsigned int)T - 1 ? s : (unsigned int)T - 1;
			tmpf[0] = static_cast<float>(s);
		}
		cutoff++;
		wsync();
		sample = (unsigned int)tmpf[0];
		wsync();
		center_color = texel4(w, (int)sample);
		if (w.lane == 0) {
			cluster_centers[(int)clusters_selected] = center_color;
		}
		clusters_selected++;
		if (clusters_selected >= partition_count) {
			break;
		}
		first = false;
	}
	wsync();
}

// kmeans_assign :146-207
ASTC_COOP void kmeans_assign(WCtx w, unsigned int partition_count, unsigned int L) {
	PartScratch ps = make_part_scratch(w, L);
	SPtr<f4> cluster_centers = centers_of(w);
	int T = w.T;
	f4 cw = bi_of(w).channel_weight;
	ASTC_NOUNROLL
	for (int i = w.lane; i < T; i += ASTC_WARP) {
		float best_distance = 3.402823466e+38f;
		unsigned int best_partition = 0;
		f4 color = texel4(w, i);
		ASTC_NOUNROLL
		for (unsigned int j = 0; j < partition_count; j++) {
			f4 diff = color - cluster_centers[(int)j];
			float distance = dot_s(diff * diff, cw);
			if (distance < best_distance) {
				best_distance = distance;
				best_partition = j;
			}
		}
		ps.pot[i] = static_cast<uint8_t>(best_partition);
	}
	wsync();
	if (w.lane == 0) {
		uint32_t cnt = 0;         // four u8 counters
		ASTC_NOUNROLL
		for (int i = 0; i < T; i++) {
			cnt += 1u << (8 * ps.pot[i]);
		}
		bool problem_case;
		do {
			problem_case = false;
			ASTC_NOUNROLL
			for (unsigned int i = 0; i < partition_count; i++) {
				if (((cnt >> (8 * i)) & 0xFF) == 0) {
					cnt -= 1u << (8 * ps.pot[(int)i]);
					cnt += 1u << (8 * i);
					ps.pot[(int)i] = static_cast<uint8_t>(i);
					problem_case = true;
				}
			}
		} while (problem_case);
	}
	wsync();
}

// kmeans_update :210-243 - one chain per (partition, channel) in texel order
ASTC_COOP void kmeans_update(WCtx w, unsigned int partition_count, unsigned int L) {
	PartScratch ps = make_part_scratch(w, L);
	SPtr<float> centers = sptr<float>(centers_of(w).off);
	int T = w.T;
	ASTC_NOUNROLL
	for (int id = w.lane; id < (int)partition_count * 4; id += ASTC_WARP) {
		unsigned int p = (unsigned int)id >> 2;
		int c = id & 3;
		SPtr<float> d = blk_of(w, c);
		float s = 0.0f;
		int n = 0;
		ASTC_NOUNROLL
		for (int i = 0; i < T; i++) {
			if (ps.pot[i] == p) {
				s = s + d[i];
				n++;
			}
		}
		float scale = 1.0f / static_cast<float>(n);
		centers[id] = s * scale;
	}
	wsync();
}

ASTC_FN int min3i(int a, int b, int c) { return mini(mini(a, b), c); }
ASTC_FN int min4i(int a, int b, int c, int d) { return mini(mini(a, b), mini(c, d)); }

ASTC_FN uint8_t partition_mismatch2(const uint64_t a[2], const uint64_t b[2]) {   // :253-263
	int v1 = ASTC_POPCLL(a[0] ^ b[0]) + ASTC_POPCLL(a[1] ^ b[1]);
	int v2 = ASTC_POPCLL(a[0] ^ b[1]) + ASTC_POPCLL(a[1] ^ b[0]);
	return static_cast<uint8_t>(mini(v1, v2) / 2);
}

ASTC_NOINLINE uint8_t partition_mismatch3(const uint64_t a[3], const uint64_t b[3]) {   // :273-304
	int p00 = ASTC_POPCLL(a[0] ^ b[0]), p01 = ASTC_POPCLL(a[0] ^ b[1]), p02 = ASTC_POPCLL(a[0] ^ b[2]);
	int p10 = ASTC_POPCLL(a[1] ^ b[0]), p11 = ASTC_POPCLL(a[1] ^ b[1]), p12 = ASTC_POPCLL(a[1] ^ b[2]);
	int p20 = ASTC_POPCLL(a[2] ^ b[0]), p21 = ASTC_POPCLL(a[2] ^ b[1]), p22 = ASTC_POPCLL(a[2] ^ b[2]);
	int v0 = mini(p11 + p22, p12 + p21) + p00;
	int v1 = mini(p10 + p22, p12 + p20) + p01;
	int v2 = mini(p10 + p21, p11 + p20) + p02;
	return static_cast<uint8_t>(min3i(v0, v1, v2) / 2);
}

ASTC_NOINLINE uint8_t partition_mismatch4(const uint64_t a[4], const uint64_t b[4]) {   // :314-353
	int p00 = ASTC_POPCLL(a[0] ^ b[0]), p01 = ASTC_POPCLL(a[0] ^ b[1]), p02 = ASTC_POPCLL(a[0] ^ b[2]), p03 = ASTC_POPCLL(a[0] ^ b[3]);
	int p10 = ASTC_POPCLL(a[1] ^ b[0]), p11 = ASTC_POPCLL(a[1] ^ b[1]), p12 = ASTC_POPCLL(a[1] ^ b[2]), p13 = ASTC_POPCLL(a[1] ^ b[3]);
	int p20 = ASTC_POPCLL(a[2] ^ b[0]), p21 = ASTC_POPCLL(a[2] ^ b[1]), p22 = ASTC_POPCLL(a[2] ^ b[2]), p23 = ASTC_POPCLL(a[2] ^ b[3]);
	int p30 = ASTC_POPCLL(a[3] ^ b[0]), p31 = ASTC_POPCLL(a[3] ^ b[1]), p32 = ASTC_POPCLL(a[3] ^ b[2]), p33 = ASTC_POPCLL(a[3] ^ b[3]);
	int mx23 = mini(p22 + p33, p23 + p32);
	int mx13 = mini(p21 + p33, p23 + p31);
	int mx12 = mini(p21 + p32, p22 + p31);
	int mx03 = mini(p20 + p33, p23 + p30);
	int mx02 = mini(p20 + p32, p22 + p30);
	int mx01 = mini(p21 + p30, p20 + p31);
	int v0 = p00 + min3i(p11 + mx23, p12 + mx13, p13 + mx12);
	int v1 = p01 + min3i(p10 + mx23, p12 + mx03, p13 + mx02);
	int v2 = p02 + min3i(p11 + mx03, p10 + mx13, p13 + mx01);
	int v3 = p03 + min3i(p11 + mx02, p12 + mx01, p10 + mx12);
	return static_cast<uint8_t>(min4i(v0, v1, v2, v3) / 2);
}

// compute_kmeans_partition_ordering :458-509; only the first L entries of the ordering are materialised.
ASTC_COOP unsigned int compute_kmeans_partition_ordering(WCtx w, unsigned int partition_count, unsigned int L) {
	PartScratch ps = make_part_scratch(w, L);
	for (unsigned int i = 0; i < 3; i++) {
		if (i == 0) {
			kmeans_init(w, partition_count, L);
		} else {
			kmeans_update(w, partition_count, L);
		}
		kmeans_assign(w, partition_count, L);
	}
	unsigned int texels_to_process = (unsigned int)w.T < ASTC_MAX_KMEANS_TEXELS ? (unsigned int)w.T : (unsigned int)ASTC_MAX_KMEANS_TEXELS;
	for (unsigned int p = (unsigned int)w.lane; p < 4; p += ASTC_WARP) {
		uint64_t bm = 0;
		if (p < partition_count) {
			ASTC_NOUNROLL
			for (unsigned int i = 0; i < texels_to_process; i++) {
				if (ps.pot[BSD.kmeans_texels[i]] == p) {
					bm |= 1ULL << i;
				}
			}
		}
		ps.bitmaps[(int)p] = bm;
	}
	for (int i = w.lane; i < 64; i += ASTC_WARP) {
		ps.hist[i] = 0;
	}
	wsync();
	uint64_t bitmaps[4] = {ps.bitmaps[0], ps.bitmaps[1], ps.bitmaps[2], ps.bitmaps[3]};
	unsigned int active_count = BSD.partitioning_count_selected[partition_count - 1];
	const uint64_t* cov = BSD.coverage_bitmaps[partition_count];
	ASTC_NOUNROLL
	for (unsigned int i = (unsigned int)w.lane; i < active_count; i += ASTC_WARP) {
		uint64_t b[4];
		for (unsigned int k = 0; k < partition_count; k++) {
			b[k] = ASTC_LDG(cov + (size_t)i * partition_count + k);
		}
		uint8_t m;
		if (partition_count == 2) m = partition_mismatch2(bitmaps, b);
		else if (partition_count == 3) m = partition_mismatch3(bitmaps, b);
		else m = partition_mismatch4(bitmaps, b);
		ps.mism[(int)i] = m;
	}
	wsync();
	// stable counting sort (:412-455): histogram ...
	unsigned int rounds = (active_count + ASTC_WARP - 1) / ASTC_WARP;
	ASTC_NOUNROLL
	for (unsigned int r = 0; r < rounds; r++) {
		unsigned int i = r * ASTC_WARP + (unsigned int)w.lane;
		int key = i < active_count ? ps.mism[(int)i] : 255;
		int rank = wsame_key_rank(key, w.lane);
		int cnt = wsame_key_count(key);
		if (rank == 0 && key < 64) {
			ps.hist[key] = (uint16_t)(ps.hist[key] + cnt);
		}
		wsync();
	}
	// ... exclusive prefix over the mismatch values ...
	if (w.lane == 0) {
		uint16_t sum = 0;
		ASTC_NOUNROLL
		for (unsigned int i = 0; i < texels_to_process; i++) {
			uint16_t c = ps.hist[(int)i];
			ps.hist[(int)i] = sum;
			sum = (uint16_t)(sum + c);
		}
	}
	wsync();
	// ... and placement in index order
	ASTC_NOUNROLL
	for (unsigned int r = 0; r < rounds; r++) {
		unsigned int i = r * ASTC_WARP + (unsigned int)w.lane;
		int key = i < active_count ? ps.mism[(int)i] : 255;
		int rank = wsame_key_rank(key, w.lane);
		int cnt = wsame_key_count(key);
		unsigned int pos = 0xFFFFFFFFu;
		if (key < 64) {
			pos = (unsigned int)ps.hist[key] + (unsigned int)rank;
			if (pos < L) {
				ps.order[(int)pos] = (uint16_t)i;
			}
		}
		wsync();
		if (rank == cnt - 1 && key < 64) {
			ps.hist[key] = (uint16_t)(ps.hist[key] + cnt);
		}
		wsync();
	}
	return active_count;
}

// Evaluate one candidate partitioning on one lane: compute_avgs_and_dirs_{4_comp,3_comp_rgb}
// (averages_and_directions.cpp:388-456, :568-628) + compute_error_squared_{rgba,rgb} (:723-945) + the
// line-length penalty of find_best_partition_candidates (:676-690, :733-747).
ASTC_NOINLINE void evaluate_partitioning(WCtx w, unsigned int pc, unsigned int packed, bool uses_alpha, float weight_imprecision_estim,
                                   float& uncor_error_out, float& samec_error_out) {
	PartView pi = part_view_packed(pc, packed);
	int T = w.T;
	int ncomp = uses_alpha ? 4 : 3;
	f4 dmean = bi_of(w).data_mean;
	f4 mean = uses_alpha ? dmean : mk4(dmean.x, dmean.y, dmean.z, 0.0f);
	SPtr<float> b0 = blk_of(w, 0);
	uint32_t cs = tp4(w);
	f4 averages[4];
	{
		f4 block_total = mean * static_cast<float>(T);
		f4 rest = block_total;
		ASTC_NOUNROLL
		for (unsigned int p = 0; p < pc - 1; p++) {
			f4 total = splat4(0.0f);
			ASTC_NOUNROLL
			for (int c = 0; c < ncomp; c++) {
				SPtr<float> d = sptr<float>(b0.off + (uint32_t)c * cs);
				float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
				int i = 0;
				ASTC_NOUNROLL
				for (; i + 3 < T; i += 4) {
					if (pi.partition_of_texel[i] == p) a0 = a0 + d[i];
					if (pi.partition_of_texel[i + 1] == p) a1 = a1 + d[i + 1];
					if (pi.partition_of_texel[i + 2] == p) a2 = a2 + d[i + 2];
					if (pi.partition_of_texel[i + 3] == p) a3 = a3 + d[i + 3];
				}
				if (i < T && pi.partition_of_texel[i] == p) a0 = a0 + d[i];
				if (i + 1 < T && pi.partition_of_texel[i + 1] == p) a1 = a1 + d[i + 1];
				if (i + 2 < T && pi.partition_of_texel[i + 2] == p) a2 = a2 + d[i + 2];
				set_lane(total, c, (a0 + a2) + (a1 + a3));
			}
			rest = rest - total;
			averages[p] = total / static_cast<float>(pv_count(pi, p));
		}
		averages[pc - 1] = rest / static_cast<float>(pv_count(pi, pc - 1));
	}
	float uacc0 = 0.0f, uacc1 = 0.0f, uacc2 = 0.0f, uacc3 = 0.0f;
	float sacc0 = 0.0f, sacc1 = 0.0f, sacc2 = 0.0f, sacc3 = 0.0f;
	float penalty_u[4], penalty_s[4];
	f4 ew = bi_of(w).channel_weight;
	ASTC_NOUNROLL
	for (unsigned int p = 0; p < pc; p++) {
		const uint8_t* tix = pi.texels + pv_start(pi, p);
		int n = pv_count(pi, p);
		f4 average = averages[p];
		f4 sum_xp = splat4(0.0f), sum_yp = splat4(0.0f), sum_zp = splat4(0.0f), sum_wp = splat4(0.0f);
		ASTC_NOUNROLL
		for (int i = 0; i < n; i++) {
			int t = tix[i];
			SPtr<float> tx = b0 + t;
			f4 d = mk4(tx[0], sptr<float>(tx.off + cs)[0], sptr<float>(tx.off + 2 * cs)[0], uses_alpha ? sptr<float>(tx.off + 3 * cs)[0] : 0.0f);
			d = d - average;
			f4 zero = splat4(0.0f);
			sum_xp = sum_xp + (d.x > 0.0f ? d : zero);
			sum_yp = sum_yp + (d.y > 0.0f ? d : zero);
			sum_zp = sum_zp + (d.z > 0.0f ? d : zero);
			if (uses_alpha) {
				sum_wp = sum_wp + (d.w > 0.0f ? d : zero);
			}
		}
		f4 best_vector = sum_xp;
		float best_sum = dot_s(sum_xp, sum_xp);
		float prod_yp = dot_s(sum_yp, sum_yp);
		if (prod_yp > best_sum) {
			best_vector = sum_yp;
			best_sum = prod_yp;
		}
		float prod_zp = dot_s(sum_zp, sum_zp);
		if (prod_zp > best_sum) {
			best_vector = sum_zp;
			best_sum = prod_zp;
		}
		if (uses_alpha) {
			float prod_wp = dot_s(sum_wp, sum_wp);
			if (prod_wp > best_sum) {
				best_vector = sum_wp;
			}
		}
		f4 ub = normalize_safe4(best_vector, uses_alpha ? unit4() : unit3());
		f4 sb = normalize_safe4(average, uses_alpha ? unit4() : unit3());
		f4 ua = uses_alpha ? average - ub * splat4(dot_s(average, ub)) : average - ub * dot3_splat(average, ub);
		float lo = 1e10f, hi = -1e10f;
		ASTC_NOUNROLL
		for (int i = 0; i < n; i++) {
			int t = tix[i];
			SPtr<float> tx = b0 + t;
			float r = tx[0], g = sptr<float>(tx.off + cs)[0], b = sptr<float>(tx.off + 2 * cs)[0];
			float uparam, uerr, serr;
			if (uses_alpha) {
				float a = sptr<float>(tx.off + 3 * cs)[0];
				uparam = (r * ub.x) + (g * ub.y) + (b * ub.z) + (a * ub.w);
				float d0 = (ua.x - r) + (uparam * ub.x);
				float d1 = (ua.y - g) + (uparam * ub.y);
				float d2 = (ua.z - b) + (uparam * ub.z);
				float d3 = (ua.w - a) + (uparam * ub.w);
				uerr = (ew.x * d0 * d0) + (ew.y * d1 * d1) + (ew.z * d2 * d2) + (ew.w * d3 * d3);
				float sparam = (r * sb.x) + (g * sb.y) + (b * sb.z) + (a * sb.w);
				float s0 = sparam * sb.x - r;
				float s1 = sparam * sb.y - g;
				float s2 = sparam * sb.z - b;
				float s3 = sparam * sb.w - a;
				serr = (ew.x * s0 * s0) + (ew.y * s1 * s1) + (ew.z * s2 * s2) + (ew.w * s3 * s3);
			} else {
				uparam = (r * ub.x) + (g * ub.y) + (b * ub.z);
				float d0 = (ua.x - r) + (uparam * ub.x);
				float d1 = (ua.y - g) + (uparam * ub.y);
				float d2 = (ua.z - b) + (uparam * ub.z);
				uerr = (ew.x * d0 * d0) + (ew.y * d1 * d1) + (ew.z * d2 * d2);
				float sparam = (r * sb.x) + (g * sb.y) + (b * sb.z);
				float s0 = sparam * sb.x - r;
				float s1 = sparam * sb.y - g;
				float s2 = sparam * sb.z - b;
				serr = (ew.x * s0 * s0) + (ew.y * s1 * s1) + (ew.z * s2 * s2);
			}
			lo = minf(uparam, lo);
			hi = maxf(uparam, hi);
			int l = i & 3;
			if (l == 0) { uacc0 = uacc0 + uerr; sacc0 = sacc0 + serr; }
			else if (l == 1) { uacc1 = uacc1 + uerr; sacc1 = sacc1 + serr; }
			else if (l == 2) { uacc2 = uacc2 + uerr; sacc2 = sacc2 + serr; }
			else { uacc3 = uacc3 + uerr; sacc3 = sacc3 + serr; }
		}
		float line_length = maxf(hi - lo, 1e-7f);
		float tpp = static_cast<float>(n);
		f4 error_weights = splat4(tpp * weight_imprecision_estim);
		f4 uncor_vector = ub * line_length;
		f4 samec_vector = sb * line_length;
		if (uses_alpha) {
			penalty_u[p] = dot_s(uncor_vector * uncor_vector, error_weights);
			penalty_s[p] = dot_s(samec_vector * samec_vector, error_weights);
		} else {
			penalty_u[p] = dot3_s(uncor_vector * uncor_vector, error_weights);
			penalty_s[p] = dot3_s(samec_vector * samec_vector, error_weights);
		}
	}
	float uncor_error = (uacc0 + uacc2) + (uacc1 + uacc3);
	float samec_error = (sacc0 + sacc2) + (sacc1 + sacc3);
	for (unsigned int p = 0; p < pc; p++) {
		uncor_error += penalty_u[p];
		samec_error += penalty_s[p];
	}
	uncor_error_out = uncor_error;
	samec_error_out = samec_error;
}

ASTC_FN void insert_result(unsigned int max_values, float this_error, unsigned int this_partition, float* best_errors, unsigned int* best_partitions) {   // :512-548
	if (this_error >= best_errors[max_values - 1]) {
		return;
	}
	for (unsigned int i = 0; i < max_values; i++) {
		if (this_error > best_errors[i]) {
			continue;
		}
		for (unsigned int j = max_values - 1; j > i; j--) {
			best_errors[j] = best_errors[j - 1];
			best_partitions[j] = best_partitions[j - 1];
		}
		best_errors[i] = this_error;
		best_partitions[i] = this_partition;
		break;
	}
}

// find_best_partition_candidates :551-780
ASTC_COOP unsigned int find_best_partition_candidates(WCtx w, unsigned int partition_count, unsigned int partition_search_limit,
                                                      unsigned int best_partitions[8], unsigned int requested_candidates) {
	unsigned int texels_per_block = (unsigned int)w.T;
	float weight_imprecision_estim = 0.055f;
	if (texels_per_block <= 20) weight_imprecision_estim = 0.03f;
	else if (texels_per_block <= 31) weight_imprecision_estim = 0.04f;
	else if (texels_per_block <= 41) weight_imprecision_estim = 0.05f;
	weight_imprecision_estim = weight_imprecision_estim * weight_imprecision_estim;

	unsigned int n = BSD.partitioning_count_selected[partition_count - 1];
	unsigned int L = partition_search_limit < n ? partition_search_limit : n;
	PartScratch ps = make_part_scratch(w, L);
	unsigned int sequence_len = compute_kmeans_partition_ordering(w, partition_count, L);
	partition_search_limit = partition_search_limit < sequence_len ? partition_search_limit : sequence_len;
	requested_candidates = partition_search_limit < requested_candidates ? partition_search_limit : requested_candidates;
	bool uses_alpha = !is_constant_channel(w, 3);

	ASTC_NOUNROLL
	for (unsigned int i = (unsigned int)w.lane; i < partition_search_limit; i += ASTC_WARP) {
		float ue, se;
		evaluate_partitioning(w, partition_count, ps.order[(int)i], uses_alpha, weight_imprecision_estim, ue, se);
		ps.cand_err[2 * (int)i] = ue;
		ps.cand_err[2 * (int)i + 1] = se;
	}
	wsync();
	float uncor_best_errors[8], samec_best_errors[8];
	unsigned int uncor_best_partitions[8], samec_best_partitions[8];
	for (unsigned int i = 0; i < 8; i++) {
		uncor_best_partitions[i] = 0;
		samec_best_partitions[i] = 0;
		uncor_best_errors[i] = ERROR_CALC_DEFAULT;
		samec_best_errors[i] = ERROR_CALC_DEFAULT;
	}
	ASTC_NOUNROLL
	for (unsigned int i = 0; i < partition_search_limit; i++) {
		unsigned int partition = ps.order[(int)i];
		insert_result(requested_candidates, ps.cand_err[2 * (int)i], partition, uncor_best_errors, uncor_best_partitions);
		insert_result(requested_candidates, ps.cand_err[2 * (int)i + 1], partition, samec_best_errors, samec_best_partitions);
	}
	unsigned int interleave[16];
	for (unsigned int i = 0; i < requested_candidates; i++) {
		interleave[2 * i] = part_view_packed(partition_count, uncor_best_partitions[i]).partition_index;
		interleave[2 * i + 1] = part_view_packed(partition_count, samec_best_partitions[i]).partition_index;
	}
	unsigned int emitted = 0;
	for (unsigned int i = 0; i < requested_candidates * 2; i++) {
		unsigned int partition = interleave[i];
		bool written = false;
		for (unsigned int k = 0; k < emitted; k++) {
			written = written || best_partitions[k] == partition;
		}
		if (!written) {
			best_partitions[emitted] = partition;
			emitted++;
			if (emitted == requested_candidates) {
				break;
			}
		}
	}
	wsync();
	return emitted;
}
