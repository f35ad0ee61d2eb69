version = 'MMult_cuda_11';
GPU Device 0: "NVIDIA B200" with compute capability 10.0

MY_MMult = [

 error: i 0  j 0 diff 6.143436  got -19.135733  expect -25.279169 diff too big !
