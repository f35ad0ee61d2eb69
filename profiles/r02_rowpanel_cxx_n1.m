version = 'b200gemm_rowpanel_cxx';
% b200gemm 0.2 (sm_100a; tcgen05+TMA; round 2); M = 1 x 4096 rows, N = 4096, K = 4096, B broadcast from rank 0 inside every call
MY_MMult = [
1 434440.68 4.017696e-06 
];
