"""Host-side mirror of the reference's AES services (same class / method names, argument
meaning and error behaviour), so code written against songhayeong/aes-fhe keeps working
and so the hot path can be benchmarked on a box where /root/reference does not exist."""
