#define LIBAV_VERSION "oracle-refbuild"
