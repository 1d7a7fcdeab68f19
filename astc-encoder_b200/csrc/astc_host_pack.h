// Packs the host-built tables into the compact device layout (astc_dev_tables.h) and plans the
// per-warp arena. The packed image is one relocatable byte blob: pointer fields of DevBsd first hold
// byte offsets into the blob, relocate_bsd() turns them into addresses for a given base (a device
// allocation for the product, the host blob itself for the host-side kernel simulator used in tests).
#pragma once
#include <vector>
#include <cstring>
#include <cmath>
#include "astc_dev_tables.h"
#include "astc_host_tables.h"

namespace astc_host {

struct PackedTables {
	std::vector<uint8_t> blob;
	DevBsd bsd;              // pointer fields = offsets until relocate_bsd()
	DevBsd bsd_1p;           // the same tables with the compact one-plane arena plan (DevBsd::layout_planes == 1)
	DevConstTables consts;
};

static inline size_t blob_append(std::vector<uint8_t>& blob, const void* data, size_t bytes, size_t align = 16) {
	size_t off = (blob.size() + align - 1) / align * align;
	blob.resize(off + bytes);
	if (data) {
		memcpy(blob.data() + off, data, bytes);
	}
	return off;
}

static inline void fill_dev_const_tables(DevConstTables& d) {
	const ConstTables& ct = const_tables();
	memset(&d, 0, sizeof(d));
	for (int t4 = 0; t4 < 3; t4++)
		for (int t3 = 0; t3 < 3; t3++)
			for (int t2 = 0; t2 < 3; t2++)
				for (int t1 = 0; t1 < 3; t1++)
					for (int t0 = 0; t0 < 3; t0++)
						d.integer_of_trits[(((t4 * 3 + t3) * 3 + t2) * 3 + t1) * 3 + t0] = ct.integer_of_trits[t4][t3][t2][t1][t0];
	for (int q2 = 0; q2 < 5; q2++)
		for (int q1 = 0; q1 < 5; q1++)
			for (int q0 = 0; q0 < 5; q0++)
				d.integer_of_quints[(q2 * 5 + q1) * 5 + q0] = ct.integer_of_quints[q2][q1][q0];
	memcpy(d.color_unquant_to_uquant, ct.color_unquant_to_uquant, sizeof(d.color_unquant_to_uquant));
	memcpy(d.color_uquant_to_scrambled_pquant, ct.color_uquant_to_scrambled_pquant, sizeof(d.color_uquant_to_scrambled_pquant));
	memcpy(d.quant_mode_table, ct.quant_mode_table, sizeof(d.quant_mode_table));
	memcpy(d.trits_of_integer, ct.trits_of_integer, sizeof(d.trits_of_integer));
	memcpy(d.quints_of_integer, ct.quints_of_integer, sizeof(d.quints_of_integer));
	memcpy(d.color_scrambled_pquant_to_uquant, ct.color_scrambled_pquant_to_uquant, sizeof(d.color_scrambled_pquant_to_uquant));
	for (int q = 0; q < 12; q++) {
		memcpy(d.wq_quant_to_unquant[q], ct.weight_quant[q].quant_to_unquant, 32);
		memcpy(d.wq_scramble_map[q], ct.weight_quant[q].scramble_map, 32);
		memcpy(d.wq_prev_next[q], ct.weight_quant[q].prev_next_values, sizeof(uint16_t) * 65);
		memcpy(d.wq_unscramble_and_unquant[q], ct.weight_quant[q].unscramble_and_unquant_map, 32);
	}
	for (int j = 0; j < 64; j++) {
		for (int i = 0; i < ASTC_ANGULAR_STEPS; i++) {
			d.sin_table[j][i] = ct.sin_table[j][i];
			d.cos_table[j][i] = ct.cos_table[j][i];
		}
	}
}

static inline uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

// partition_index_limit[3] / decides the scratch needed by the partition search
static inline void pack_device_tables(const BlockSizeTables& t, const unsigned int partition_index_limit[3], PackedTables& out) {
	std::vector<uint8_t>& blob = out.blob;
	blob.clear();
	DevBsd& b = out.bsd;
	memset(&b, 0, sizeof(b));
	b.dim_x = t.dim_x;
	b.dim_y = t.dim_y;
	b.dim_z = t.dim_z;
	b.texel_count = t.texel_count;
	b.decimation_mode_count_always = t.decimation_mode_count_always;
	b.decimation_mode_count_selected = t.decimation_mode_count_selected;
	b.decimation_mode_count_all = t.decimation_mode_count_all;
	b.block_mode_count_1plane_always = t.block_mode_count_1plane_always;
	b.block_mode_count_1plane_selected = t.block_mode_count_1plane_selected;
	b.block_mode_count_1plane_2plane_selected = t.block_mode_count_1plane_2plane_selected;
	b.block_mode_count_all = t.block_mode_count_all;
	for (int i = 0; i < 4; i++) {
		b.partitioning_count_selected[i] = t.partitioning_count_selected[i];
	}
	memcpy(b.kmeans_texels, t.kmeans_texels, sizeof(b.kmeans_texels));
	const unsigned int T = t.texel_count;

	// block modes
	std::vector<DevBlockMode> bms(t.block_mode_count_all);
	for (unsigned int i = 0; i < t.block_mode_count_all; i++) {
		bms[i].mode_index = t.block_modes[i].mode_index;
		bms[i].decimation_mode = t.block_modes[i].decimation_mode;
		bms[i].quant_mode = t.block_modes[i].quant_mode;
		bms[i].weight_bits = t.block_modes[i].weight_bits;
		bms[i].is_dual_plane = t.block_modes[i].is_dual_plane;
	}
	size_t off_bm = blob_append(blob, bms.data(), bms.size() * sizeof(DevBlockMode));
	size_t off_bmpi = blob_append(blob, t.block_mode_packed_index, sizeof(t.block_mode_packed_index));

	// decimation modes + blobs
	std::vector<DevDecMode> dms(t.decimation_mode_count_all);
	std::vector<uint8_t> dblob;
	uint32_t dwi_total = 0, dwi_total_1p = 0;
	unsigned int max_wtc = 1;
	for (unsigned int d = 0; d < t.decimation_mode_count_all; d++) {
		const DecimationInfo& di = t.decimation_tables[d];
		DevDecMode& dm = dms[d];
		dm.maxprec_1plane = t.decimation_modes[d].maxprec_1plane;
		dm.maxprec_2planes = t.decimation_modes[d].maxprec_2planes;
		dm.refprec_1plane = t.decimation_modes[d].refprec_1plane;
		dm.refprec_2planes = t.decimation_modes[d].refprec_2planes;
		dm.weight_count = di.weight_count;
		dm.weight_x = di.weight_x;
		dm.weight_y = di.weight_y;
		dm.weight_z = di.weight_z;
		dm.max_texel_weight_count = di.max_texel_weight_count;
		const unsigned int W = di.weight_count;
		const unsigned int E = di.weight_texel_offset[W];
		size_t base = (dblob.size() + 15) / 16 * 16;
		size_t twi_off = 16 * (size_t)T;
		size_t tci_off = 20 * (size_t)T;
		size_t wto_off = 24 * (size_t)T;
		size_t wtc_off = wto_off + 2 * (W + 1);
		size_t total = wtc_off + 2 * (size_t)E;
		dblob.resize(base + total, 0);
		uint8_t* p = dblob.data() + base;
		float* tcf = reinterpret_cast<float*>(p);
		uint32_t* twi = reinterpret_cast<uint32_t*>(p + twi_off);
		uint32_t* tci = reinterpret_cast<uint32_t*>(p + tci_off);
		for (unsigned int i = 0; i < T; i++) {
			uint32_t wi = 0, ci = 0;
			for (unsigned int k = 0; k < 4; k++) {
				uint32_t c = di.texel_weight_contribs_int[k][i];
				wi |= (uint32_t)di.texel_weights[k][i] << (8 * k);
				ci |= c << (8 * k);
				tcf[4 * i + k] = static_cast<float>(c) * (1.0f / 16.0f);
			}
			twi[i] = wi;
			tci[i] = ci;
		}
		uint16_t* wto = reinterpret_cast<uint16_t*>(p + wto_off);
		for (unsigned int i = 0; i <= W; i++) {
			wto[i] = di.weight_texel_offset[i];
		}
		uint16_t* wtc = reinterpret_cast<uint16_t*>(p + wtc_off);
		for (unsigned int e = 0; e < E; e++) {
			wtc[e] = (uint16_t)(di.weight_texels[e] | ((unsigned int)di.weight_texel_contribs[e] << 8));
		}
		unsigned int grid_max_wtc = 0;
		for (unsigned int i = 0; i < W; i++) {
			if (di.weight_texel_count[i] > max_wtc) {
				max_wtc = di.weight_texel_count[i];
			}
			if (di.weight_texel_count[i] > grid_max_wtc) {
				grid_max_wtc = di.weight_texel_count[i];
			}
		}
		dm.blob_offset = (uint32_t)base;
		dm.wto_offset = (uint16_t)wto_off;
		dm.wtc_offset = (uint16_t)wtc_off;
		dm.max_weight_texels = (uint16_t)grid_max_wtc;
		// arena slot for the decimated ideal weights (only grids the search can reference)
		dm.dwi_offset = (uint16_t)dwi_total;
		dm.dwi_offset_1p = (uint16_t)dwi_total_1p;
		if (d < t.decimation_mode_count_selected) {
			b.dec_stage_bytes = (uint32_t)((dblob.size() + 15) / 16 * 16);
			dwi_total += W * (dm.maxprec_2planes >= 0 ? 2u : 1u);
			dwi_total = (dwi_total + 3u) & ~3u;
			dwi_total_1p += W;
			dwi_total_1p = (dwi_total_1p + 3u) & ~3u;
		}
	}
	b.max_weight_texel_count = (uint8_t)max_wtc;
	size_t off_dm = blob_append(blob, dms.data(), dms.size() * sizeof(DevDecMode));
	size_t off_dblob = blob_append(blob, dblob.data(), dblob.size());

	// partitions
	uint32_t stride = (ASTC_PART_HDR + 2 * T + 3u) & ~3u;
	b.part_stride = stride;
	size_t off_part[5] = {0, 0, 0, 0, 0};
	size_t off_ppi[3] = {0, 0, 0};
	size_t off_cov[5] = {0, 0, 0, 0, 0};
	for (unsigned int pc = 1; pc <= 4; pc++) {
		unsigned int n = pc == 1 ? 1 : t.partitioning_count_all[pc - 1];
		std::vector<uint8_t> pb((size_t)(n ? n : 1) * stride, 0);
		for (unsigned int i = 0; i < n; i++) {
			const PartitionInfo& pi = t.partitionings[pc][i];
			uint8_t* e = pb.data() + (size_t)i * stride;
			e[0] = (uint8_t)(pi.partition_index & 0xFF);
			e[1] = (uint8_t)(pi.partition_index >> 8);
			unsigned int k = 0;
			for (unsigned int p = 0; p < 4; p++) {
				e[2 + p] = pi.partition_texel_count[p];
				for (unsigned int j = 0; j < pi.partition_texel_count[p]; j++) {
					e[ASTC_PART_HDR + T + k++] = pi.texels_of_partition[p][j];
				}
			}
			memcpy(e + ASTC_PART_HDR, pi.partition_of_texel, T);
		}
		off_part[pc] = blob_append(blob, pb.data(), pb.size());
		if (pc >= 2) {
			off_ppi[pc - 2] = blob_append(blob, t.partitioning_packed_index[pc - 2], sizeof(uint16_t) * ASTC_MAX_PARTITIONINGS);
			unsigned int ns = t.partitioning_count_selected[pc - 1];
			off_cov[pc] = blob_append(blob, t.coverage_bitmaps[pc], sizeof(uint64_t) * (size_t)(ns ? ns : 1) * pc);
		}
	}
	// store offsets in the pointer fields
	b.block_modes = reinterpret_cast<const DevBlockMode*>(off_bm);
	b.block_mode_packed_index = reinterpret_cast<const uint16_t*>(off_bmpi);
	b.dec_modes = reinterpret_cast<const DevDecMode*>(off_dm);
	b.dec_blob = reinterpret_cast<const uint8_t*>(off_dblob);
	for (unsigned int pc = 1; pc <= 4; pc++) {
		b.partitions[pc] = reinterpret_cast<const uint8_t*>(off_part[pc]);
		if (pc >= 2) {
			b.partitioning_packed_index[pc - 2] = reinterpret_cast<const uint16_t*>(off_ppi[pc - 2]);
			b.coverage_bitmaps[pc] = reinterpret_cast<const uint64_t*>(off_cov[pc]);
		}
	}

	// ---- per-warp arena plan (fixed head: see the A_* constants in astc_dev_core.cuh) ----
	const uint32_t Tp = (T + 3u) & ~3u;
	uint32_t o = ASTC_ARENA_FIXED;                                              // A_BLK
	o = align16(o + 16 * Tp);                                                   // block texels [4][Tp]
	b.record_bytes = align16(ASTC_ARENA_PERSIST_HEAD + 16 * Tp);
	b.off_scratch = o;
	// union scratch: the largest of the phase layouts (see astc_dev_search.cuh / astc_dev_partition.cuh)
	uint32_t su = 32 * 68;                                                      // quantise+score rows
	uint32_t ef = 4 * 21 * 4 * 4 + 21 * 13 * 4 + 4 * 21 * 4 + 21 * 13 * 4;      // EfTables
	if (ef > su) su = ef;
	uint32_t rf = 4 * 3 * Tp + 256 + 20 * 33 * 4 + 2 * Tp;                      // RefineScratch incl. the chain tile
	if (rf > su) su = rf;
	uint32_t inf = 8 * Tp;                                                      // infilled[2][T]
	if (inf > su) su = inf;
	uint32_t ang = align16(dwi_total) + 288 + 25 * 96;                          // angular search: >= 96 step records per round
	if (ang > su) su = ang;
	for (unsigned int pc = 2; pc <= 4; pc++) {
		unsigned int n = t.partitioning_count_selected[pc - 1];
		unsigned int L = partition_index_limit[pc - 2] < n ? partition_index_limit[pc - 2] : n;
		uint32_t ps = 4 * Tp + 8 * L + 8 + 32 + 2 * ((L + 1) & ~1u) + 128 + Tp + n + 16;
		if (ps > su) su = ps;
	}
	b.scratch_bytes = align16(su);
	o = align16(b.off_scratch + b.scratch_bytes);
	b.arena_bytes_small = o;
	b.off_ei = o;         o = align16(o + 16 * Tp);                             // ideal weights / error scales, 2 planes
	b.off_dwi = o;        o = align16(o + 4 * (dwi_total ? dwi_total : 4));
	b.off_lowhigh = o;    o = align16(o + 128 * t.decimation_mode_count_selected);
	b.off_mode_err = o;   o = align16(o + 4 * t.block_mode_count_1plane_2plane_selected);
	b.arena_bytes = o;
	b.layout_planes = 2;
	// the compact plan for one-plane trials: same head, texels and scratch (the records and the refinement kernels never see
	// the tail), one plane's worth of ideal weights, decimated weights and angular ranges, errors of the 1-plane modes only
	DevBsd& c = out.bsd_1p;
	c = b;
	o = b.arena_bytes_small;
	c.off_ei = o;         o = align16(o + 8 * Tp);
	c.off_dwi = o;        o = align16(o + 4 * (dwi_total_1p ? dwi_total_1p : 4));
	c.off_lowhigh = o;    o = align16(o + 64 * t.decimation_mode_count_selected);
	c.off_mode_err = o;   o = align16(o + 4 * t.block_mode_count_1plane_selected);
	c.arena_bytes = o;
	c.layout_planes = 1;

}

static inline void relocate_bsd(DevBsd& b, const uint8_t* base);
static inline void relocate_tables(PackedTables& pk, const uint8_t* base) {
	relocate_bsd(pk.bsd, base);
	relocate_bsd(pk.bsd_1p, base);
}
static inline void relocate_bsd(DevBsd& b, const uint8_t* base) {
	auto rel = [&](const void* p) { return base + reinterpret_cast<uintptr_t>(p); };
	b.block_modes = reinterpret_cast<const DevBlockMode*>(rel(b.block_modes));
	b.block_mode_packed_index = reinterpret_cast<const uint16_t*>(rel(b.block_mode_packed_index));
	b.dec_modes = reinterpret_cast<const DevDecMode*>(rel(b.dec_modes));
	b.dec_blob = rel(b.dec_blob);
	for (unsigned int pc = 1; pc <= 4; pc++) {
		b.partitions[pc] = rel(b.partitions[pc]);
		if (pc >= 2) {
			b.partitioning_packed_index[pc - 2] = reinterpret_cast<const uint16_t*>(rel(b.partitioning_packed_index[pc - 2]));
			b.coverage_bitmaps[pc] = reinterpret_cast<const uint64_t*>(rel(b.coverage_bitmaps[pc]));
		}
	}
}

}  // namespace astc_host
