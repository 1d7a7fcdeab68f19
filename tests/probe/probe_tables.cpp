// Dev-time probe (needs /root/reference; never runs on the GPU box): compares the oracle's generated
// tables with the reference's own tables, entry by entry - 2D and 3D block sizes.
//   g++ -std=c++14 -O1 -I/root/reference/Source -DASTCENC_SSE=0 -DASTCENC_AVX=0 -DASTCENC_NEON=0 -DASTCENC_POPCNT=0 -DASTCENC_F16C=0 \
//       tests/probe/probe_tables.cpp oracle/astc_tables.cpp /root/reference/Source/astcenc_{block_sizes,partition_tables,percentile_tables,quantization,weight_quant_xfer_tables,mathlib,mathlib_softfloat}.cpp -o /tmp/probe && /tmp/probe
#include "astcenc_integer_sequence.cpp"   // for the static BISE tables
// -DPROBE_PRODUCT: probe the PRODUCT's host-side table construction (astc-encoder_b200/csrc/astc_host_tables.cpp, link that file
// instead of oracle/astc_tables.cpp) - the tables the GPU kernels read are then compared with the reference's directly.
#if defined(PROBE_PRODUCT)
#include "../../astc-encoder_b200/csrc/astc_host_tables.h"
namespace ao = astc_host;
#else
#include "../../oracle/astc_tables.h"
#endif
#include <cstdio>
#include <cstring>

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails < 40) { printf("FAIL: " __VA_ARGS__); printf("\n"); } fails++; } } while (0)

static void check_bsd(unsigned bx, unsigned by, bool can_omit, unsigned pcut, float mcut, unsigned bz = 1) {
	block_size_descriptor* bsd = aligned_malloc<block_size_descriptor>(sizeof(block_size_descriptor), 64);
	init_block_size_descriptor(bx, by, bz, can_omit, pcut, mcut, *bsd);
	ao::BlockSizeTables* t = ao::build_block_size_tables(bx, by, bz, can_omit, pcut, mcut);
	CHECK(bsd->texel_count == t->texel_count, "texel_count");
	CHECK(bsd->decimation_mode_count_always == t->decimation_mode_count_always, "dm always %u %u", bsd->decimation_mode_count_always, t->decimation_mode_count_always);
	CHECK(bsd->decimation_mode_count_selected == t->decimation_mode_count_selected, "dm sel");
	CHECK(bsd->decimation_mode_count_all == t->decimation_mode_count_all, "dm all");
	CHECK(bsd->block_mode_count_1plane_always == t->block_mode_count_1plane_always, "bm always");
	CHECK(bsd->block_mode_count_1plane_selected == t->block_mode_count_1plane_selected, "bm 1p sel");
	CHECK(bsd->block_mode_count_1plane_2plane_selected == t->block_mode_count_1plane_2plane_selected, "bm 2p sel");
	CHECK(bsd->block_mode_count_all == t->block_mode_count_all, "bm all %u %u", bsd->block_mode_count_all, t->block_mode_count_all);
	for (unsigned i = 0; i < 2048; i++) CHECK(bsd->block_mode_packed_index[i] == t->block_mode_packed_index[i], "bm packed idx %u", i);
	for (unsigned i = 0; i < bsd->block_mode_count_all; i++) {
		const block_mode& a = bsd->block_modes[i];
		const ao::BlockMode& b = t->block_modes[i];
		CHECK(a.mode_index == b.mode_index && a.decimation_mode == b.decimation_mode && a.quant_mode == b.quant_mode &&
		      a.weight_bits == b.weight_bits && a.is_dual_plane == b.is_dual_plane, "block mode %u", i);
	}
	for (unsigned i = 0; i < bsd->decimation_mode_count_all; i++) {
		const decimation_mode& a = bsd->decimation_modes[i];
		const ao::DecimationMode& b = t->decimation_modes[i];
		CHECK(a.maxprec_1plane == b.maxprec_1plane && a.maxprec_2planes == b.maxprec_2planes &&
		      a.refprec_1plane == b.refprec_1plane && a.refprec_2planes == b.refprec_2planes, "dec mode %u", i);
		const decimation_info& da = bsd->decimation_tables[i];
		const ao::DecimationInfo& db = t->decimation_tables[i];
		CHECK(da.texel_count == db.texel_count && da.weight_count == db.weight_count && da.weight_x == db.weight_x &&
		      da.weight_y == db.weight_y && da.weight_z == db.weight_z && da.max_texel_weight_count == db.max_texel_weight_count, "di hdr %u", i);
		for (unsigned tix = 0; tix < da.texel_count; tix++) {
			CHECK(da.texel_weight_count[tix] == db.texel_weight_count[tix], "di twc");
			for (int k = 0; k < 4; k++) {
				CHECK(da.texel_weights_tr[k][tix] == db.texel_weights[k][tix], "di tw %u %u %d", i, tix, k);
				CHECK(da.texel_weight_contribs_int_tr[k][tix] == db.texel_weight_contribs_int[k][tix], "di twi");
				CHECK(da.texel_weight_contribs_float_tr[k][tix] == db.texel_weight_contribs_float[k][tix], "di twf");
			}
		}
		for (unsigned w = 0; w < da.weight_count; w++) {
			CHECK(da.weight_texel_count[w] == db.weight_texel_count[w], "di wtc");
			for (unsigned j = 0; j < da.weight_texel_count[w]; j++) {
				unsigned o = db.weight_texel_offset[w] + j;
				CHECK(da.weight_texels_tr[j][w] == db.weight_texels[o], "di wt");
				CHECK(da.weights_texel_contribs_tr[j][w] == db.weight_texel_contribs[o], "di wtcf");
				CHECK(da.texel_contrib_for_weight[j][w] == db.texel_contrib_for_weight[o], "di tcfw");
			}
		}
	}
	for (int i = 0; i < 64 && i < bsd->texel_count; i++) CHECK(bsd->kmeans_texels[i] == t->kmeans_texels[i], "kmeans texel %d", i);
	for (unsigned pc = 1; pc <= 4; pc++) {
		CHECK(bsd->partitioning_count_selected[pc - 1] == t->partitioning_count_selected[pc - 1], "pcount sel %u: %u %u", pc, bsd->partitioning_count_selected[pc - 1], t->partitioning_count_selected[pc - 1]);
		CHECK(bsd->partitioning_count_all[pc - 1] == t->partitioning_count_all[pc - 1], "pcount all %u", pc);
		if (pc >= 2) for (unsigned i = 0; i < 1024; i++) CHECK(bsd->partitioning_packed_index[pc - 2][i] == t->partitioning_packed_index[pc - 2][i], "ppacked");
		for (unsigned i = 0; i < bsd->partitioning_count_all[pc - 1]; i++) {
			const partition_info& a = bsd->get_raw_partition_info(pc, i);
			const ao::PartitionInfo& b = t->partitionings[pc][i];
			CHECK(a.partition_count == b.partition_count && a.partition_index == b.partition_index, "pi hdr");
			for (unsigned p = 0; p < 4; p++) {
				CHECK(a.partition_texel_count[p] == b.partition_texel_count[p], "pi ptc");
				for (unsigned j = 0; j < a.partition_texel_count[p]; j++) CHECK(a.texels_of_partition[p][j] == b.texels_of_partition[p][j], "pi top");
			}
			for (unsigned tix = 0; tix < bsd->texel_count; tix++) CHECK(a.partition_of_texel[tix] == b.partition_of_texel[tix], "pi pot");
			if (pc >= 2 && i < bsd->partitioning_count_selected[pc - 1]) {
				const uint64_t* ca = pc == 2 ? bsd->coverage_bitmaps_2[i] : pc == 3 ? bsd->coverage_bitmaps_3[i] : bsd->coverage_bitmaps_4[i];
				for (unsigned p = 0; p < pc; p++) CHECK(ca[p] == t->coverage_bitmaps[pc][i * pc + p], "coverage");
			}
		}
	}
	printf("bsd %ux%ux%u omit=%d pcut=%u mcut=%.2f: modes %u/%u/%u/%u dec %u/%u/%u parts %u/%u/%u  fails so far %d\n", bx, by, bz, can_omit, pcut, mcut,
	       t->block_mode_count_1plane_always, t->block_mode_count_1plane_selected, t->block_mode_count_1plane_2plane_selected, t->block_mode_count_all,
	       t->decimation_mode_count_always, t->decimation_mode_count_selected, t->decimation_mode_count_all,
	       t->partitioning_count_selected[1], t->partitioning_count_selected[2], t->partitioning_count_selected[3], fails);
	ao::free_block_size_tables(t);
	aligned_free<block_size_descriptor>(bsd);
}

int main() {
	const ao::ConstTables& ct = ao::const_tables();
	CHECK(memcmp(ct.trits_of_integer, trits_of_integer, sizeof(trits_of_integer)) == 0, "trits_of_integer");
	CHECK(memcmp(ct.quints_of_integer, quints_of_integer, sizeof(quints_of_integer)) == 0, "quints_of_integer");
	CHECK(memcmp(ct.integer_of_trits, integer_of_trits, sizeof(integer_of_trits)) == 0, "integer_of_trits");
	CHECK(memcmp(ct.integer_of_quints, integer_of_quints, sizeof(integer_of_quints)) == 0, "integer_of_quints");
	for (int q = 0; q <= 20; q++) {
		unsigned b, t, qn; ao::ise_btq(q, b, t, qn);
		CHECK(b == btq_counts[q].bits && t == btq_counts[q].trits && qn == btq_counts[q].quints, "btq %d", q);
		for (unsigned n = 1; n <= 64; n++) CHECK(ao::ise_sequence_bitcount(n, q) == get_ise_sequence_bitcount(n, (quant_method)q), "bitcount %d %u", q, n);
		CHECK(ao::get_quant_level(q) == get_quant_level((quant_method)q), "level");
	}
	CHECK(memcmp(ct.color_unquant_to_uquant, color_unquant_to_uquant_tables, sizeof(color_unquant_to_uquant_tables)) == 0, "color_unquant_to_uquant");
	for (int qi = 0; qi < 17; qi++) {
		unsigned levels = get_quant_level((quant_method)(QUANT_6 + qi));
		for (unsigned p = 0; p < levels; p++) CHECK(ct.color_scrambled_pquant_to_uquant[qi][p] == color_scrambled_pquant_to_uquant_tables[qi][p], "pq2uq %d %u", qi, p);
		for (unsigned p = 0; p < levels; p++) {
			unsigned u = color_scrambled_pquant_to_uquant_tables[qi][p];
			CHECK(ct.color_uquant_to_scrambled_pquant[qi][u] == color_uquant_to_scrambled_pquant_tables[qi][u], "uq2pq(level) %d %u", qi, u);
		}
		int diff = 0;
		for (unsigned u = 0; u < 256; u++) diff += ct.color_uquant_to_scrambled_pquant[qi][u] != color_uquant_to_scrambled_pquant_tables[qi][u];
		if (diff) printf("note: uq2pq[%d] differs at %d non-level inputs\n", qi, diff);
	}
	CHECK(memcmp(ct.quant_mode_table, quant_mode_table, sizeof(quant_mode_table)) == 0, "quant_mode_table");
	for (int q = 0; q < 12; q++) {
		unsigned levels = get_quant_level((quant_method)q);
		const quant_and_transfer_table& r = quant_and_xfer_tables[q];
		for (unsigned i = 0; i < levels; i++) {
			CHECK(r.quant_to_unquant[i] == ct.weight_quant[q].quant_to_unquant[i], "w q2u %d %u", q, i);
			CHECK(r.scramble_map[i] == ct.weight_quant[q].scramble_map[i], "w scr %d %u", q, i);
			CHECK(r.unscramble_and_unquant_map[i] == ct.weight_quant[q].unscramble_and_unquant_map[i], "w unscr %d %u", q, i);
			unsigned v = r.quant_to_unquant[i];
			CHECK(r.prev_next_values[v] == ct.weight_quant[q].prev_next_values[v], "w pn %d %u: %x %x", q, v, r.prev_next_values[v], ct.weight_quant[q].prev_next_values[v]);
		}
	}
	printf("const tables done, fails %d\n", fails);
	check_bsd(4, 4, true, 3, 0.55f);
	check_bsd(6, 6, true, 3, 0.77f);
	check_bsd(6, 6, false, 3, 0.77f);
	check_bsd(8, 8, true, 4, 0.93f);
	check_bsd(5, 4, true, 2, 0.43f);
	check_bsd(10, 8, true, 4, 1.0f);
	check_bsd(12, 12, true, 4, 0.98f);
	check_bsd(12, 10, false, 2, 0.40f);
	// the ten 3D footprints (no percentile selection: every grid that fits, every legal mode)
	static const unsigned fp3[10][3] = {{3, 3, 3}, {4, 3, 3}, {4, 4, 3}, {4, 4, 4}, {5, 4, 4}, {5, 5, 4}, {5, 5, 5}, {6, 5, 5}, {6, 6, 5}, {6, 6, 6}};
	for (int i = 0; i < 10; i++) {
		check_bsd(fp3[i][0], fp3[i][1], (i & 1) != 0, 2 + (unsigned)(i % 3), 0.77f, fp3[i][2]);
	}
	printf("TOTAL FAILS %d\n", fails);
	return fails != 0;
}
