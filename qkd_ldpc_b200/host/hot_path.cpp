// The reference's per-frame hot-path functions as batch-of-one calls into libqkdldpc_b200.so.
//   calculate_syndrome_regular / _irregular   <- reference src/array_and_matrix_operations.cpp:463-486
//   sum_product_decoding_regular / _irregular <- reference src/qkd_ldpc_algorithm.cpp:3-173 / 175-345
//   QKD_LDPC_regular / _irregular             <- reference src/qkd_ldpc_algorithm.cpp:347-396 / 398-447
// The *_regular variants of the reference loop over max_*_weight instead of the per-node weights; on a matrix whose
// header weights are consistent the two are the same function, and an inconsistent "regular" matrix is rejected here
// rather than read out of bounds.
#include <stdexcept>

#include "device_bridge.hpp"
#include "qkd_ldpc.hpp"

namespace
{
    void require_consistent_regular(const H_matrix &h)
    {
        for (size_t i = 0; i < h.num_bit_nodes; ++i)
            if (static_cast<size_t>(h.bit_nodes_weight[i]) != h.max_bit_nodes_weight)
                throw std::runtime_error("regular decoding requested on a matrix whose bit weights differ from max_bit_nodes_weight");
        for (size_t j = 0; j < h.num_check_nodes; ++j)
            if (static_cast<size_t>(h.check_nodes_weight[j]) != h.max_check_nodes_weight)
                throw std::runtime_error("regular decoding requested on a matrix whose check weights differ from max_check_nodes_weight");
    }

    void syndrome(const int *bits, const H_matrix &h, int *out)
    {
        qkd_b200::check(qlb_syndrome_batch(qkd_b200::context(), qkd_b200::code_for(h), 1, bits, out), "qlb_syndrome_batch");
    }

    SP_result decode(const double *llr, const H_matrix &h, const int *syndrome_in, size_t max_it, double thr, int *bits_out)
    {
        const qlb_decode_params p = qkd_b200::params_from_cfg(max_it, thr);
        uint32_t iterations = 0;
        uint8_t result = 0;
        qkd_b200::check(qlb_sum_product_batch(qkd_b200::context(), qkd_b200::code_for(h), &p, 1, llr, syndrome_in, bits_out, &iterations, &result),
                        "qlb_sum_product_batch");
        return {iterations, (result & QLB_RES_SYNDROMES_MATCH) != 0};
    }

    LDPC_result reconcile(const int *alice, const int *bob, double qber, const H_matrix &h)
    {
        const qlb_decode_params p = qkd_b200::params_from_cfg(CFG.SUM_PRODUCT_MAX_ITERATIONS, CFG.SUM_PRODUCT_MSG_LLR_THRESHOLD);
        uint32_t iterations = 0;
        uint8_t result = 0;
        qkd_b200::check(qlb_reconcile_batch(qkd_b200::context(), qkd_b200::code_for(h), &p, 1, alice, bob, &qber, &iterations, &result, nullptr, nullptr),
                        "qlb_reconcile_batch");
        LDPC_result r;
        r.sp_res = {iterations, (result & QLB_RES_SYNDROMES_MATCH) != 0};
        r.keys_match = (result & QLB_RES_KEYS_MATCH) != 0;
        return r;
    }
}

void calculate_syndrome_regular(const int *const bit_array, const H_matrix &matrix, int *const syndrome_out)
{
    require_consistent_regular(matrix);
    syndrome(bit_array, matrix, syndrome_out);
}

void calculate_syndrome_irregular(const int *const bit_array, const H_matrix &matrix, int *const syndrome_out)
{
    syndrome(bit_array, matrix, syndrome_out);
}

SP_result sum_product_decoding_regular(const double *const bit_array_llr, const H_matrix &matrix, const int *const syndrome,
                                       const size_t &max_num_iterations, const double &msg_threshold, int *const bit_array_out)
{
    require_consistent_regular(matrix);
    return decode(bit_array_llr, matrix, syndrome, max_num_iterations, msg_threshold, bit_array_out);
}

SP_result sum_product_decoding_irregular(const double *const bit_array_llr, const H_matrix &matrix, const int *const syndrome,
                                         const size_t &max_num_iterations, const double &msg_threshold, int *const bit_array_out)
{
    return decode(bit_array_llr, matrix, syndrome, max_num_iterations, msg_threshold, bit_array_out);
}

LDPC_result QKD_LDPC_regular(const int *const alice_bit_array, const int *const bob_bit_array, const double &QBER, const H_matrix &matrix)
{
    require_consistent_regular(matrix);
    return reconcile(alice_bit_array, bob_bit_array, QBER, matrix);
}

LDPC_result QKD_LDPC_irregular(const int *const alice_bit_array, const int *const bob_bit_array, const double &QBER, const H_matrix &matrix)
{
    return reconcile(alice_bit_array, bob_bit_array, QBER, matrix);
}
