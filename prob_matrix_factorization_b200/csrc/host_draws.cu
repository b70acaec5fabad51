// a2 (host side): NumPy's Generator(PCG64).gamma(1.0, scale) stream, replayed bit for bit by several host threads.
//
// The reference draws its initial state with np.random.default_rng(seed).gamma(1.0, 0.1, size=(R, K)) (poisson_mf_cavi.py:62-63,
// hpf_cavi.py:71-80): (2M + 500k) x 64 x 2 = 320 M variates at BASELINE config C5, 3.1 s of a single sequential NumPy stream --
// 75 % of Model(config).fit(DataFrame) once the sweeps run on a B200.  Parity needs the SAME numbers, so the stream is replayed,
// not replaced:
//   * PCG64 (128-bit LCG, XSL-RR output) can jump ahead in O(log n), so any raw position is a cheap starting point;
//   * shape == 1.0 makes gamma() the ziggurat exponential: 98.9 % of the variates consume one raw draw, the rest two (a second
//     draw decides tail / accept / reject-and-retry), so "which raw positions start an attempt" is a chain that depends on
//     everything before it -- but two chains that start one position apart merge after a few positions.
// Each thread therefore walks its block of raw positions under both hypotheses for its first position (an attempt starts
// there / it is the second draw of the previous block's last attempt), a short sequential pass picks the true hypothesis per
// block and the output offsets, and a second parallel pass writes the variates in place.  Ziggurat tables: the 256-entry
// ke / we / fe tables of NumPy's exponential ziggurat (Marsaglia & Tsang; values as compiled into numpy/random/_generator).
// tests/test_host_draws_cpu.py compares against the installed NumPy bit for bit (values and final generator state).
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pmf {
namespace {

typedef unsigned __int128 u128;
constexpr double kZigExpR = 7.69711747013104972;

const uint64_t kKe[256] = {
    7971545857431494ULL, 0ULL, 5485857970336126ULL, 6877400373607440ULL,
    7489560515621038ULL, 7829793950745724ULL, 8045251395085594ULL, 8193552821270898ULL,
    8301707212298418ULL, 8384003209374832ULL, 8448689755168200ULL, 8500854585063478ULL,
    8543802742323106ULL, 8579772857648236ULL, 8610334328270398ULL, 8636619566280862ULL,
    8659465946817878ULL, 8679505875409358ULL, 8697225801520776ULL, 8713005977443536ULL,
    8727147906454692ULL, 8739893704890038ULL, 8751440024696698ULL, 8761948238062960ULL,
    8771552003860596ULL, 8780362968290610ULL, 8788475114930438ULL, 8795968123070796ULL,
    8802909988292858ULL, 8809359087581710ULL, 8815365821575970ULL, 8820973931588800ULL,
    8826221564107158ULL, 8831142137483404ULL, 8835765052397426ULL, 8840116277974648ULL,
    8844218838221542ULL, 8848093218006260ULL, 8851757703688506ULL, 8855228670347734ULL,
    8858520825126080ULL, 8861647414312948ULL, 8864620400320394ULL, 8867450613535032ULL,
    8870147883110754ULL, 8872721150032146ULL, 8875178565190242ULL, 8877527574738170ULL,
    8879774994610604ULL, 8881927075778634ULL, 8883989561556502ULL, 8885967738067168ULL,
    8887866478800856ULL, 8889690284057818ULL, 8891443315947666ULL, 8893129429518482ULL,
    8894752200505984ULL, 8896314950123264ULL, 8897820767252858ULL, 8899272528353282ULL,
    8900672915349966ULL, 8902024431744704ULL, 8903329417147192ULL, 8904590060406012ULL,
    8905808411494018ULL, 8906986392283810ULL, 8908125806332284ULL, 8909228347778946ULL,
    8910295609450180ULL, 8911329090250868ULL, 8912330201915374ULL, 8913300275181656ULL,
    8914240565445170ULL, 8915152257942916ULL, 8916036472512488ULL, 8916894267966144ULL,
    8917726646115692ULL, 8918534555480190ULL, 8919318894705170ULL, 8920080515719156ULL,
    8920820226650618ULL, 8921538794526266ULL, 8922236947769418ULL, 8922915378515480ULL,
    8923574744759820ULL, 8924215672351958ULL, 8924838756848636ULL, 8925444565237162ULL,
    8926033637539416ULL, 8926606488305930ULL, 8927163608008600ULL, 8927705464339880ULL,
    8928232503425544ULL, 8928745150957558ULL, 8929243813252980ULL, 8929728878244356ULL,
    8930200716406566ULL, 8930659681624710ULL, 8931106112007190ULL, 8931540330647876ULL,
    8931962646340834ULL, 8932373354250910ULL, 8932772736543124ULL, 8933161062973652ULL,
    8933538591444894ULL, 8933905568527004ULL, 8934262229948010ULL, 8934608801054528ULL,
    8934945497244894ULL, 8935272524376414ULL, 8935590079148328ULL, 8935898349461876ULL,
    8936197514758882ULL, 8936487746340036ULL, 8936769207664048ULL, 8937042054628744ULL,
    8937306435835058ULL, 8937562492834854ULL, 8937810360363402ULL, 8938050166557284ULL,
    8938282033158466ULL, 8938506075705162ULL, 8938722403710132ULL, 8938931120826946ULL,
    8939132325004766ULL, 8939326108632062ULL, 8939512558669762ULL, 8939691756774158ULL,
    8939863779409990ULL, 8940028697953972ULL, 8940186578789100ULL, 8940337483389966ULL,
    8940481468399302ULL, 8940618585695992ULL, 8940748882454662ULL, 8940872401197050ULL,
    8940989179835208ULL, 8941099251706688ULL, 8941202645601704ULL, 8941299385782368ULL,
    8941389491993960ULL, 8941472979468230ULL, 8941549858918698ULL, 8941620136527848ULL,
    8941683813926148ULL, 8941740888162738ULL, 8941791351667640ULL, 8941835192205302ULL,
    8941872392819226ULL, 8941902931767472ULL, 8941926782448690ULL, 8941943913318396ULL,
    8941954287795084ULL, 8941957864155806ULL, 8941954595420724ULL, 8941944429226144ULL,
    8941927307685492ULL, 8941903167237602ULL, 8941871938481652ULL, 8941833545998016ULL,
    8941787908154234ULL, 8941734936895206ULL, 8941674537516674ULL, 8941606608420918ULL,
    8941531040853536ULL, 8941447718620056ULL, 8941356517781006ULL, 8941257306323958ULL,
    8941149943810914ULL, 8941034280999228ULL, 8940910159434164ULL, 8940777411010892ULL,
    8940635857503634ULL, 8940485310059376ULL, 8940325568653336ULL, 8940156421503112ULL,
    8939977644438114ULL, 8939789000220574ULL, 8939590237814000ULL, 8939381091594596ULL,
    8939161280500638ULL, 8938930507114326ULL, 8938688456670012ULL, 8938434795982096ULL,
    8938169172285110ULL, 8937891211977748ULL, 8937600519261604ULL, 8937296674664432ULL,
    8936979233436516ULL, 8936647723807416ULL, 8936301645088910ULL, 8935940465608202ULL,
    8935563620453574ULL, 8935170509012442ULL, 8934760492279316ULL, 8934332889908232ULL,
    8933886976981030ULL, 8933421980459034ULL, 8932937075281380ULL, 8932431380068266ULL,
    8931903952381602ULL, 8931353783488908ULL, 8930779792568492ULL, 8930180820284952ULL,
    8929555621653500ULL, 8928902858099222ULL, 8928221088602964ULL, 8927508759808348ULL,
    8926764194944404ULL, 8925985581394242ULL, 8925170956711904ULL, 8924318192855506ULL,
    8923424978364230ULL, 8922488798157886ULL, 8921506910578770ULL, 8920476321224196ULL,
    8919393753031106ULL, 8918255611967910ULL, 8917057947558148ULL, 8915796407299442ULL,
    8914466183841290ULL, 8913061953535784ULL, 8911577804662434ULL, 8910007153233212ULL,
    8908342643782164ULL, 8906576031902208ULL, 8904698044465300ULL, 8902698212389652ULL,
    8900564669414922ULL, 8898283908495804ULL, 8895840484961220ULL, 8893216652275640ULL,
    8890391911743352ULL, 8887342451323378ULL, 8884040440144924ULL, 8880453133239800ULL,
    8876541723776518ULL, 8872259855113102ULL, 8867551668208538ULL, 8862349204777254ULL,
    8856568902200012ULL, 8850106784293916ULL, 8842831740745002ULL, 8834575940248166ULL,
    8825120832349124ULL, 8814176156651890ULL, 8801347484544986ULL, 8786084197194146ULL,
    8767592496903178ULL, 8744682338845716ULL, 8715480686119910ULL, 8676850260251934ULL,
    8623083654098352ULL, 8542525795804796ULL, 8406823688997808ULL, 8122426762520768ULL};
const double kWe[256] = {
    0x1.164ec94bf5dc1p-50, 0x1.0589d8b5d4119p-57, 0x1.ad6b2495b4d2bp-57, 0x1.19335a95b8dbap-56,
    0x1.522e6e54a2a73p-56, 0x1.85090fbc27a80p-56, 0x1.b38d1ef79b7ccp-56, 0x1.decd8b76dbd98p-56,
    0x1.03bf049c65c3cp-55, 0x1.170db24d6f670p-55, 0x1.2980290da2633p-55, 0x1.3b388fe3d6ecap-55,
    0x1.4c515c60bfe21p-55, 0x1.5cdf89d024ac3p-55, 0x1.6cf40f0a72bbdp-55, 0x1.7c9cdda17d019p-55,
    0x1.8be5954d3606fp-55, 0x1.9ad80552237d2p-55, 0x1.a97c8be5d5203p-55, 0x1.b7da5dddda3c4p-55,
    0x1.c5f7bd78c3f89p-55, 0x1.d3da24df17c36p-55, 0x1.e186678f1735ap-55, 0x1.ef00ccf5f4faap-55,
    0x1.fc4d25d683209p-55, 0x1.04b76ed6a7558p-54, 0x1.0b348479b80fcp-54, 0x1.119f38749f5afp-54,
    0x1.17f8ceb4bdfa0p-54, 0x1.1e426e93e49e7p-54, 0x1.247d26538ff2ep-54, 0x1.2aa9ee123680bp-54,
    0x1.30c9aa526da4bp-54, 0x1.36dd2e26d8202p-54, 0x1.3ce53d12162a0p-54, 0x1.42e28ca706748p-54,
    0x1.48d5c5f35e712p-54, 0x1.4ebf86bcd0b93p-54, 0x1.54a0629786f4dp-54, 0x1.5a78e3db8befdp-54,
    0x1.60498c7dd2ecfp-54, 0x1.6612d6d0c68e0p-54, 0x1.6bd5362faa944p-54, 0x1.71911797990bbp-54,
    0x1.7746e23077973p-54, 0x1.7cf6f7c7e8172p-54, 0x1.82a1b53fed599p-54, 0x1.884772f2be1ecp-54,
    0x1.8de8850d0c52ap-54, 0x1.93853bdfda244p-54, 0x1.991de42ad1338p-54, 0x1.9eb2c75ff03bfp-54,
    0x1.a4442be14884ap-54, 0x1.a9d255396d261p-54, 0x1.af5d844f224c9p-54, 0x1.b4e5f794c979bp-54,
    0x1.ba6beb33f8f89p-54, 0x1.bfef99359fe99p-54, 0x1.c57139a70d29fp-54, 0x1.caf102bc25adbp-54,
    0x1.d06f28ef0e6fbp-54, 0x1.d5ebdf1d86b8dp-54, 0x1.db6756a429057p-54, 0x1.e0e1bf77c31fep-54,
    0x1.e65b483cf1044p-54, 0x1.ebd41e5e21b62p-54, 0x1.f14c6e202949fp-54, 0x1.f6c462b57feb5p-54,
    0x1.fc3c26504a9a1p-54, 0x1.00d9f119a3cd9p-53, 0x1.0395df60db162p-53, 0x1.0651f1c7276f8p-53,
    0x1.090e3bb4b0072p-53, 0x1.0bcad03710137p-53, 0x1.0e87c207a2f66p-53, 0x1.114523917ac15p-53,
    0x1.140306f707dbep-53, 0x1.16c17e1777ffbp-53, 0x1.19809a93d2396p-53, 0x1.1c406dd3d5283p-53,
    0x1.1f01090a9c4e2p-53, 0x1.21c27d3b10e05p-53, 0x1.2484db3c2a329p-53, 0x1.274833bd0189fp-53,
    0x1.2a0c9748bcdaap-53, 0x1.2cd2164a53b5dp-53, 0x1.2f98c11031721p-53, 0x1.3260a7cfb7611p-53,
    0x1.3529daa8a1ba1p-53, 0x1.37f469a851af0p-53, 0x1.3ac064ccfeffcp-53, 0x1.3d8ddc08d336dp-53,
    0x1.405cdf44f09c4p-53, 0x1.432d7e6466cd0p-53, 0x1.45ffc94716ca7p-53, 0x1.48d3cfcc883c4p-53,
    0x1.4ba9a1d6b18a4p-53, 0x1.4e814f4cb45eap-53, 0x1.515ae81d900fbp-53, 0x1.54367c42cb5f8p-53,
    0x1.57141bc316f27p-53, 0x1.59f3d6b4e9cf9p-53, 0x1.5cd5bd4119335p-53, 0x1.5fb9dfa56cf26p-53,
    0x1.62a04e3731a2ep-53, 0x1.65891965c9b8cp-53, 0x1.687451bd3ebeep-53, 0x1.6b6207e8d3cdfp-53,
    0x1.6e524cb59a608p-53, 0x1.714531150a9fbp-53, 0x1.743ac61fa041cp-53, 0x1.77331d177d130p-53,
    0x1.7a2e476b1240ap-53, 0x1.7d2c56b7d17f7p-53, 0x1.802d5ccce7277p-53, 0x1.83316badfe62ap-53,
    0x1.86389596108e7p-53, 0x1.8942ecfa40f54p-53, 0x1.8c50848cc6094p-53, 0x1.8f616f3fe1513p-53,
    0x1.9275c048e73e1p-53, 0x1.958d8b235828ap-53, 0x1.98a8e3940bbf4p-53, 0x1.9bc7ddac7035dp-53,
    0x1.9eea8dcdde951p-53, 0x1.a21108ad0592dp-53, 0x1.a53b63556c690p-53, 0x1.a869b32d0f30fp-53,
    0x1.ab9c0df81657ap-53, 0x1.aed289dcaacffp-53, 0x1.b20d3d66e8bb5p-53, 0x1.b54c3f8cf2542p-53,
    0x1.b88fa7b324fb6p-53, 0x1.bbd78db072610p-53, 0x1.bf2409d2dfd85p-53, 0x1.c27534e42e02dp-53,
    0x1.c5cb282eab1a4p-53, 0x1.c925fd82323fbp-53, 0x1.cc85cf395a56cp-53, 0x1.cfeab83ed7180p-53,
    0x1.d354d4130f2adp-53, 0x1.d6c43ed1ea3fep-53, 0x1.da391538da50ap-53, 0x1.ddb374ad2357fp-53,
    0x1.e1337b426509bp-53, 0x1.e4b947c16a452p-53, 0x1.e844f9af4237fp-53, 0x1.ebd6b154a7678p-53,
    0x1.ef6e8fc5b9168p-53, 0x1.f30cb6ea0bc7fp-53, 0x1.f6b1498515ed0p-53, 0x1.fa5c6b3efe1e5p-53,
    0x1.fe0e40add09d8p-53, 0x1.00e377af911d4p-52, 0x1.02c34ef11391bp-52, 0x1.04a6b9e9224a3p-52,
    0x1.068dccf1126dbp-52, 0x1.08789cf3aad0fp-52, 0x1.0a673f733c819p-52, 0x1.0c59ca900946fp-52,
    0x1.0e50550efcfb7p-52, 0x1.104af660befcep-52, 0x1.1249c6a92154ap-52, 0x1.144cdec6f3a2bp-52,
    0x1.1654585c404c1p-52, 0x1.18604dd6fae9ep-52, 0x1.1a70da7a27820p-52, 0x1.1c861a6782a5ap-52,
    0x1.1ea02aa9b3370p-52, 0x1.20bf293f0f4a2p-52, 0x1.22e33524fe550p-52, 0x1.250c6e6403bbap-52,
    0x1.273af61c7daa6p-52, 0x1.296eee942532bp-52, 0x1.2ba87b445db51p-52, 0x1.2de7c0e962d70p-52,
    0x1.302ce59265965p-52, 0x1.327810b2aa7d0p-52, 0x1.34c96b33bc965p-52, 0x1.37211f88ca856p-52,
    0x1.397f59c345143p-52, 0x1.3be447a8d8b83p-52, 0x1.3e5018cadded0p-52, 0x1.40c2fe9f5eeadp-52,
    0x1.433d2c9bd42f8p-52, 0x1.45bed851bc92cp-52, 0x1.4848398d39432p-52, 0x1.4ad98a75da14cp-52,
    0x1.4d7307b1cb127p-52, 0x1.5014f08b99508p-52, 0x1.52bf871acaab2p-52, 0x1.5573106f8a75ap-52,
    0x1.582fd4c1b4461p-52, 0x1.5af61fa38e107p-52, 0x1.5dc640388bd9ep-52, 0x1.60a0897081879p-52,
    0x1.63855247b2e94p-52, 0x1.6674f60c3f432p-52, 0x1.696fd4a9748eep-52, 0x1.6c7652f9a7b1ep-52,
    0x1.6f88db1f42507p-52, 0x1.72a7dce5cd218p-52, 0x1.75d3ce2bd71c3p-52, 0x1.790d2b56b71f9p-52,
    0x1.7c5477d1476d3p-52, 0x1.7faa3e96e1412p-52, 0x1.830f12cc0bec3p-52, 0x1.8683906687342p-52,
    0x1.8a085ce695babp-52, 0x1.8d9e2823b3695p-52, 0x1.9145ad2f37544p-52, 0x1.94ffb34fc2a0ep-52,
    0x1.98cd0f18d1ad8p-52, 0x1.9caea3a24d9eap-52, 0x1.a0a563e49f178p-52, 0x1.a4b2543e84c3bp-52,
    0x1.a8d68c2ad86eap-52, 0x1.ad13382d845c4p-52, 0x1.b1699c003b60ap-52, 0x1.b5db15091ea0fp-52,
    0x1.ba691d276da5ep-52, 0x1.bf154de4bef77p-52, 0x1.c3e1641c2e0a7p-52, 0x1.c8cf442c8c8f4p-52,
    0x1.cde0fecf2a97fp-52, 0x1.d318d6b2738c5p-52, 0x1.d87946fec3becp-52, 0x1.de050af4ef19fp-52,
    0x1.e3bf26e190960p-52, 0x1.e9aaf2af383c1p-52, 0x1.efcc26750ea4ap-52, 0x1.f626e9791f7a7p-52,
    0x1.fcbfe43f6c6e5p-52, 0x1.01ce2b362ec2ep-51, 0x1.056118bf58eefp-51, 0x1.091c1cdcba54ep-51,
    0x1.0d031785d48a0p-51, 0x1.111a8034392a6p-51, 0x1.156786775442ap-51, 0x1.19f03bcb3c2d6p-51,
    0x1.1ebbca0c9fa7cp-51, 0x1.23d2bb659919fp-51, 0x1.293f5ae49aaa5p-51, 0x1.2f0e38a4411f0p-51,
    0x1.354ee27ccf75ep-51, 0x1.3c14ec7c8b861p-51, 0x1.4379766e41362p-51, 0x1.4b9d7cd4751d1p-51,
    0x1.54ad83ccf73f6p-51, 0x1.5ee7ae17313d2p-51, 0x1.6aa676d4bbf72p-51, 0x1.78750d6eac62fp-51,
    0x1.8939fe6f2ed19p-51, 0x1.9e9dc0d487b85p-51, 0x1.bc39e51da71fcp-51, 0x1.ec9d9297ebb83p-51};
const double kFe[256] = {
    0x1.0000000000000p+0, 0x1.e0545e5881137p-1, 0x1.cd0a65081fff1p-1, 0x1.be5007beb7b27p-1,
    0x1.b210f0ee67f2ap-1, 0x1.a76baa562fae7p-1, 0x1.9de9715556d9bp-1, 0x1.95431c455aa39p-1,
    0x1.8d4a376d3d22fp-1, 0x1.85de87806c5b8p-1, 0x1.7ee8a2d243126p-1, 0x1.7856e9b09d47ep-1,
    0x1.721bb5ba94b63p-1, 0x1.6c2c3498418c6p-1, 0x1.667fa6d4f5c06p-1, 0x1.610edc1a7af66p-1,
    0x1.5bd3d694cac75p-1, 0x1.56c9882da8773p-1, 0x1.51eba1578899ap-1, 0x1.4d366c151f8afp-1,
    0x1.48a6afb8ee069p-1, 0x1.44399afa8e125p-1, 0x1.3fecb2bb18b80p-1, 0x1.3bbdc44e1d114p-1,
    0x1.37aada708ddd9p-1, 0x1.33b23450e6318p-1, 0x1.2fd23e345da5ep-1, 0x1.2c098b61f4f24p-1,
    0x1.2856d111132bdp-1, 0x1.24b8e228c50a3p-1, 0x1.212eaba813ec8p-1, 0x1.1db7319877b89p-1,
    0x1.1a518c71e3b25p-1, 0x1.16fce6dce6feep-1, 0x1.13b87bc33169cp-1, 0x1.108394a1cc38dp-1,
    0x1.0d5d8812b1e2bp-1, 0x1.0a45b8854d02ap-1, 0x1.073b931ee3b7dp-1, 0x1.043e8ebd26548p-1,
    0x1.014e2b160f324p-1, 0x1.fcd3dfe214576p-2, 0x1.f722d8ebfc5fap-2, 0x1.f1886d1eb424dp-2,
    0x1.ec03d4b969d90p-2, 0x1.e6945367dd351p-2, 0x1.e139375e137fcp-2, 0x1.dbf1d88a7210cp-2,
    0x1.d6bd97db9ed7ap-2, 0x1.d19bde97e1a0bp-2, 0x1.cc8c1dc40e092p-2, 0x1.c78dcd983fb60p-2,
    0x1.c2a06d00ea583p-2, 0x1.bdc3812aeeeb5p-2, 0x1.b8f6951990b88p-2, 0x1.b43939454806fp-2,
    0x1.af8b03428ef5fp-2, 0x1.aaeb8d6fdf6e5p-2, 0x1.a65a76aa30140p-2, 0x1.a1d76207521f4p-2,
    0x1.9d61f695a3792p-2, 0x1.98f9df2097ba8p-2, 0x1.949ec9f9a8110p-2, 0x1.905068c545d04p-2,
    0x1.8c0e704b75d39p-2, 0x1.87d8984bc3f8cp-2, 0x1.83ae9b5446138p-2, 0x1.7f90369b6ce59p-2,
    0x1.7b7d29dc6801ep-2, 0x1.77753735e72e3p-2, 0x1.7378230b08deap-2, 0x1.6f85b3e649e9dp-2,
    0x1.6b9db25e4e99cp-2, 0x1.67bfe8fc60d9fp-2, 0x1.63ec2424827e4p-2, 0x1.602231fef5876p-2,
    0x1.5c61e2631ee6cp-2, 0x1.58ab06c3aa9efp-2, 0x1.54fd721bda3e7p-2, 0x1.5158f8dde89f5p-2,
    0x1.4dbd70e26f91dp-2, 0x1.4a2ab158bdad3p-2, 0x1.46a092b80beefp-2, 0x1.431eeeb1841e2p-2,
    0x1.3fa5a0230a14ep-2, 0x1.3c34830abb285p-2, 0x1.38cb747b17defp-2, 0x1.356a528fcd0ddp-2,
    0x1.3210fc6312435p-2, 0x1.2ebf520394270p-2, 0x1.2b75346ae2262p-2, 0x1.2832857457629p-2,
    0x1.24f727d4776fdp-2, 0x1.21c2ff10b7effp-2, 0x1.1e95ef77b09dbp-2, 0x1.1b6fde19abc5ap-2,
    0x1.1850b0c191982p-2, 0x1.15384dee291efp-2, 0x1.12269ccba9fbap-2, 0x1.0f1b852d9a66cp-2,
    0x1.0c16ef88f5333p-2, 0x1.0918c4ee93e13p-2, 0x1.0620ef05d90d2p-2, 0x1.032f580797c2cp-2,
    0x1.0043eab93476ap-2, 0x1.fabd24cff9354p-3, 0x1.f4fe75c963e7ep-3, 0x1.ef4ba0fe8e09bp-3,
    0x1.e9a48005940f2p-3, 0x1.e408ed62f83a7p-3, 0x1.de78c48224f39p-3, 0x1.d8f3e1ae3eeb8p-3,
    0x1.d37a220b431fdp-3, 0x1.ce0b638f6d09fp-3, 0x1.c8a784fce1802p-3, 0x1.c34e65db9afeep-3,
    0x1.bdffe67394435p-3, 0x1.b8bbe7c72e4a5p-3, 0x1.b3824b8dcef3ep-3, 0x1.ae52f42eb5b0bp-3,
    0x1.a92dc4bc03c49p-3, 0x1.a412a0edf5cbcp-3, 0x1.9f016d1e4c512p-3, 0x1.99fa0e43e1623p-3,
    0x1.94fc69ee692a1p-3, 0x1.900866425bb79p-3, 0x1.8b1de9f5062d5p-3, 0x1.863cdc48c1af9p-3,
    0x1.816525094e7e6p-3, 0x1.7c96ac8851baep-3, 0x1.77d15b99f46fep-3, 0x1.73151b91a2839p-3,
    0x1.6e61d63ee84eap-3, 0x1.69b775ea6da28p-3, 0x1.6515e5530d1acp-3, 0x1.607d0fab06a31p-3,
    0x1.5bece0954c2b6p-3, 0x1.57654422e78f5p-3, 0x1.52e626d078c49p-3, 0x1.4e6f7583cb6fap-3,
    0x1.4a011d8983096p-3, 0x1.459b0c92dccc6p-3, 0x1.413d30b386a9ap-3, 0x1.3ce7785f8a905p-3,
    0x1.3899d2694d5c9p-3, 0x1.34542dffa0cafp-3, 0x1.30167aabe7d6ep-3, 0x1.2be0a8504cf34p-3,
    0x1.27b2a72609940p-3, 0x1.238c67bbbe878p-3, 0x1.1f6ddaf3dca65p-3, 0x1.1b56f2031d666p-3,
    0x1.17479e6f0ae78p-3, 0x1.133fd20c9712fp-3, 0x1.0f3f7efec1720p-3, 0x1.0b4697b54b62fp-3,
    0x1.07550eeb7a5bep-3, 0x1.036ad7a6e7f04p-3, 0x1.ff0fca6cbea8dp-4, 0x1.f758566190414p-4,
    0x1.efaf3ae83c33cp-4, 0x1.e8146048eb9ccp-4, 0x1.e087af561bafbp-4, 0x1.d909116ad9398p-4,
    0x1.d198706914dd7p-4, 0x1.ca35b6b80fd57p-4, 0x1.c2e0cf42e10afp-4, 0x1.bb99a5771268fp-4,
    0x1.b460254356548p-4, 0x1.ad343b1655465p-4, 0x1.a615d3dd938b7p-4, 0x1.9f04dd046f428p-4,
    0x1.9801447336b70p-4, 0x1.910af88e574b9p-4, 0x1.8a21e835a533bp-4, 0x1.834602c3bc4bap-4,
    0x1.7c77380d7a6f3p-4, 0x1.75b5786193c1ep-4, 0x1.6f00b488416b6p-4, 0x1.6858ddc30b620p-4,
    0x1.61bde5ccadef7p-4, 0x1.5b2fbed91bb3ep-4, 0x1.54ae5b959d036p-4, 0x1.4e39af290d929p-4,
    0x1.47d1ad343985cp-4, 0x1.417649d25b10ep-4, 0x1.3b277999b9f9ep-4, 0x1.34e5319c6e718p-4,
    0x1.2eaf676948dd1p-4, 0x1.2886110ce0570p-4, 0x1.22692512c9d8cp-4, 0x1.1c589a86fa340p-4,
    0x1.165468f755392p-4, 0x1.105c88756ca50p-4, 0x1.0a70f19871b3bp-4, 0x1.04919d7f5c817p-4,
    0x1.fd7d0ba699676p-5, 0x1.f1ef49944e834p-5, 0x1.e679ea52eb2e5p-5, 0x1.db1ce49315810p-5,
    0x1.cfd83031e794ap-5, 0x1.c4abc640721e9p-5, 0x1.b997a10bed985p-5, 0x1.ae9bbc26a8084p-5,
    0x1.a3b81471bf138p-5, 0x1.98eca827b7c4cp-5, 0x1.8e3976e80776dp-5, 0x1.839e81c3a396bp-5,
    0x1.791bcb4ab089ep-5, 0x1.6eb1579b6af52p-5, 0x1.645f2c726a041p-5, 0x1.5a25513c5d2cap-5,
    0x1.5003cf296c5ebp-5, 0x1.45fab14266b19p-5, 0x1.3c0a047ff18ffp-5, 0x1.3231d7e3f14aep-5,
    0x1.28723c956c00cp-5, 0x1.1ecb45ff312d4p-5, 0x1.153d09f19b3a1p-5, 0x1.0bc7a0c7cd651p-5,
    0x1.026b2590dfaeep-5, 0x1.f24f6c7af9890p-6, 0x1.dffae7a517468p-6, 0x1.cdd9054331b0cp-6,
    0x1.bbea150fa5870p-6, 0x1.aa2e6e6924e9bp-6, 0x1.98a670f132a48p-6, 0x1.8752853ec9967p-6,
    0x1.76331da87fc96p-6, 0x1.6548b72a24077p-6, 0x1.5493da6ab0251p-6, 0x1.44151ce87f0bep-6,
    0x1.33cd225315d84p-6, 0x1.23bc9e1b93a32p-6, 0x1.13e4554725f5fp-6, 0x1.04452091e02f0p-6,
    0x1.e9bfdde89c7cep-7, 0x1.cb6b9146e2757p-7, 0x1.ad8fa5542c92dp-7, 0x1.902ea688fa7bdp-7,
    0x1.734b6e6aa74f5p-7, 0x1.56e930be416cbp-7, 0x1.3b0b8c1516f62p-7, 0x1.1fb69edb37671p-7,
    0x1.04ef2295fd7f9p-7, 0x1.d5751fa745dc5p-8, 0x1.a23e9d4974836p-8, 0x1.7049f37ec3620p-8,
    0x1.3fa97cee322fdp-8, 0x1.1073d69574043p-8, 0x1.c58b381cd4b11p-9, 0x1.6d888f3a1feffp-9,
    0x1.1946ba8e1a324p-9, 0x1.92bb5540c3e25p-10, 0x1.fb20af78dfcb9p-11, 0x1.dc31c329f0b4bp-12};

const u128 kMult = ((u128)2549297995355413924ULL << 64) | (u128)4865540595714422341ULL;

struct Pcg64 {
    u128 state, inc;
    inline uint64_t next64() {
        state = state * kMult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
    void advance(u128 delta) {   // state after `delta` steps (Brown, "Random number generation with arbitrary strides")
        u128 acc_mult = 1, acc_plus = 0, cur_mult = kMult, cur_plus = inc;
        while (delta > 0) {
            if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta >>= 1;
        }
        state = acc_mult * state + acc_plus;
    }
};

inline double to_double(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

// One ATTEMPT of random_standard_exponential starting with raw draw r0 (legacy-distributions.c / distributions.c):
// returns the number of raw draws consumed (1 or 2) and whether a variate came out (a rejected attempt retries from scratch).
struct Attempt { int consumed; bool accepted; double value; };
template <typename NextRaw>
inline Attempt attempt(uint64_t r0, NextRaw&& next_raw) {
    uint64_t ri = r0 >> 3;
    const unsigned idx = (unsigned)(ri & 0xFF);
    ri >>= 8;
    const double x = (double)ri * kWe[idx];
    if (ri < kKe[idx]) return {1, true, x};
    const double u = to_double(next_raw());
    if (idx == 0) return {2, true, kZigExpR - log1p(-u)};
    if ((kFe[idx - 1] - kFe[idx]) * u + kFe[idx] < exp(-x)) return {2, true, x};
    return {2, false, 0.0};
}

// Walk the attempts that START inside [begin, end) of the raw stream, the first one at `start` (begin or begin + 1).
// out != nullptr: write offset + scale * variate for output indices [out_base, ...) below out_limit.  Returns the number of
// variates produced; *next_start = raw position of the first attempt after the block (end or end + 1); *stop_pos (if the
// limit was hit) = raw position right after the last draw of the variate that reached the limit.
struct Walk { int64_t produced; int64_t next_start; int64_t stop_pos; };
Walk walk_block(const Pcg64& origin, int64_t begin, int64_t end, int64_t start, double scale, double offset, double* out,
                int64_t out_base, int64_t out_limit) {
    Pcg64 g = origin;
    g.advance((u128)start);
    int64_t pos = start, produced = 0;
    Walk w = {0, end, -1};
    while (pos < end) {
        const uint64_t r0 = g.next64();
        const Attempt a = attempt(r0, [&]() { return g.next64(); });
        pos += a.consumed;
        if (a.accepted) {
            if (out != nullptr) {
                if (out_base + produced < out_limit) {
                    volatile double t = scale * a.value;   // two roundings, as NumPy: scale * x, then offset + (.) -- never an FMA
                    out[out_base + produced] = offset + t;
                }
                if (out_base + produced == out_limit - 1) { w.stop_pos = pos; }
            }
            ++produced;
            if (out != nullptr && out_base + produced >= out_limit) break;
        }
    }
    w.produced = produced;
    w.next_start = pos < end ? end : pos;   // (pos < end only when the output limit stopped the walk)
    return w;
}

// Pass 1 for one block: variates and exit position under BOTH start hypotheses (an attempt starts at `begin` / at
// `begin + 1`), with one walk: the two chains are advanced in lockstep -- always the one that is behind -- until they reach
// the same position (a few steps: 98.9 % of the attempts consume one draw), from where they are the same chain.
void count_block(const Pcg64& origin, int64_t begin, int64_t end, int64_t cnt[2], int64_t next[2]) {
    Pcg64 g0 = origin, g1 = origin;
    g0.advance((u128)begin);
    g1.advance((u128)begin + 1);
    int64_t p0 = begin, p1 = begin + 1, c0 = 0, c1 = 0;
    while (p0 != p1 && (p0 < end || p1 < end)) {
        if (p0 < p1) {
            if (p0 >= end) break;
            Pcg64 t = g0;                       // an attempt may need the draw after its own: peek with a copy
            const uint64_t r0 = t.next64();
            const Attempt a = attempt(r0, [&]() { return t.next64(); });
            g0.advance((u128)a.consumed);
            p0 += a.consumed;
            c0 += a.accepted;
        } else {
            if (p1 >= end) break;
            Pcg64 t = g1;
            const uint64_t r0 = t.next64();
            const Attempt a = attempt(r0, [&]() { return t.next64(); });
            g1.advance((u128)a.consumed);
            p1 += a.consumed;
            c1 += a.accepted;
        }
    }
    if (p0 == p1) {                             // merged: one common walk for the rest of the block
        int64_t pos = p0, c = 0;
        while (pos < end) {
            const uint64_t r0 = g0.next64();
            const Attempt a = attempt(r0, [&]() { return g0.next64(); });
            pos += a.consumed;
            c += a.accepted;
        }
        cnt[0] = c0 + c; cnt[1] = c1 + c;
        next[0] = next[1] = pos;
        return;
    }
    // never merged inside the block (astronomically unlikely): finish the two chains separately
    const Walk w0 = walk_block(origin, begin, end, p0 < end ? p0 : end, 0.0, 0.0, nullptr, 0, 0);
    const Walk w1 = walk_block(origin, begin, end, p1 < end ? p1 : end, 0.0, 0.0, nullptr, 0, 0);
    cnt[0] = c0 + (p0 < end ? w0.produced : 0); next[0] = p0 < end ? w0.next_start : p0;
    cnt[1] = c1 + (p1 < end ? w1.produced : 0); next[1] = p1 < end ? w1.next_start : p1;
}

}  // namespace
}  // namespace pmf

using namespace pmf;

extern "C" int pmf_numpy_exponential_fill(const uint64_t* state_hi_lo, const uint64_t* inc_hi_lo, double scale, double offset,
                                          int64_t n, double* h_out, int32_t threads, uint64_t* new_state_hi_lo) {
    PMF_REQUIRE(state_hi_lo && inc_hi_lo && new_state_hi_lo && n >= 0 && (n == 0 || h_out), "bad argument");
    Pcg64 origin;
    origin.state = ((u128)state_hi_lo[0] << 64) | state_hi_lo[1];
    origin.inc = ((u128)inc_hi_lo[0] << 64) | inc_hi_lo[1];
    if (threads <= 0) threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    int64_t done = 0, raw_pos = 0;   // variates written so far; raw position of the next attempt
    while (done < n) {
        const int64_t need = n - done;
        // raw range of this round: ~1.2 % more draws than variates, split into blocks; start hypothesis of block 0 is known
        const int64_t span = need + need / 64 + 4096;
        const int64_t block = std::max<int64_t>(1 << 14, (span + (int64_t)threads * 8 - 1) / ((int64_t)threads * 8));
        const int64_t n_blocks = (span + block - 1) / block;
        const int64_t base = raw_pos;
        struct Info { int64_t cnt[2]; int64_t next[2]; };
        std::vector<Info> info((size_t)n_blocks);
        auto for_blocks = [&](auto&& fn) {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t)
                pool.emplace_back([&, t]() { for (int64_t b = t; b < n_blocks; b += threads) fn(b); });
            for (auto& th : pool) th.join();
        };
        // pass 1: per block, variates and exit position under both start hypotheses
        for_blocks([&](int64_t b) {
            const int64_t lo = base + b * block, hi = std::min(base + (b + 1) * block, base + span);
            count_block(origin, lo, hi, info[b].cnt, info[b].next);
        });
        // sequential: true hypothesis and output offset of every block
        std::vector<int64_t> off((size_t)n_blocks), hyp((size_t)n_blocks);
        int64_t o = done, h = 0, last_block = n_blocks - 1;
        for (int64_t b = 0; b < n_blocks; ++b) {
            const int64_t hi = std::min(base + (b + 1) * block, base + span);
            off[b] = o; hyp[b] = h;
            o += info[b].cnt[h];
            h = info[b].next[h] - hi;     // 0: next block starts an attempt at its first position, 1: one position later
            if (o >= n) { last_block = b; break; }
        }
        // pass 2: write
        std::vector<int64_t> stop((size_t)n_blocks, -1);
        const int64_t nb2 = last_block + 1;
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t)
                pool.emplace_back([&, t]() {
                    for (int64_t b = t; b < nb2; b += threads) {
                        const int64_t lo = base + b * block, hi = std::min(base + (b + 1) * block, base + span);
                        const Walk w = walk_block(origin, lo, hi, lo + hyp[b], scale, offset, h_out, off[b], n);
                        stop[b] = w.stop_pos;
                    }
                });
            for (auto& th : pool) th.join();
        }
        if (o >= n) {
            raw_pos = stop[last_block];
            done = n;
        } else {        // the slack was not enough (cannot happen in practice): continue after the last block
            done = o;
            raw_pos = base + span + h;
        }
    }
    Pcg64 fin = origin;
    fin.advance((u128)raw_pos);
    new_state_hi_lo[0] = (uint64_t)(fin.state >> 64);
    new_state_hi_lo[1] = (uint64_t)fin.state;
    return PMF_OK;
}

// ---- multi-threaded host conversions of the reference's input dtypes -----------------------------------------------
// A DataFrame hands fit() int64 ids and float64 ratings (load_data.py:93-105), NumPy state is float64; the engine wants
// int32 / float32.  Single-threaded NumPy astype + min + max over the 100 M-rating config cost ~0.6 s of fit(DataFrame).
namespace {
template <typename Fn>
void parallel_ranges(int64_t n, int threads, Fn&& fn) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<int64_t>(threads, std::max<int64_t>(1, n / (1 << 16)));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() { fn(t, n * t / threads, n * (t + 1) / threads); });
    for (auto& th : pool) th.join();
}
}  // namespace

extern "C" int pmf_host_i64_to_i32(const int64_t* h_in, int64_t n, int32_t* h_out, int64_t* h_min, int64_t* h_max, int32_t threads) {
    PMF_REQUIRE(n >= 0 && (n == 0 || (h_in && h_out)) && h_min && h_max, "bad argument");
    std::vector<int64_t> lo(256, INT64_MAX), hi(256, INT64_MIN);
    if (threads > 256) threads = 256;
    parallel_ranges(n, threads, [&](int t, int64_t a, int64_t b) {
        int64_t mn = INT64_MAX, mx = INT64_MIN;
        for (int64_t k = a; k < b; ++k) {
            const int64_t v = h_in[k];
            mn = v < mn ? v : mn; mx = v > mx ? v : mx;
            h_out[k] = (int32_t)v;
        }
        lo[t] = mn; hi[t] = mx;
    });
    *h_min = *std::min_element(lo.begin(), lo.end());
    *h_max = *std::max_element(hi.begin(), hi.end());
    return PMF_OK;
}

extern "C" int pmf_host_f64_to_f32(const double* h_in, int64_t n, float* h_out, int32_t threads) {
    PMF_REQUIRE(n >= 0 && (n == 0 || (h_in && h_out)), "bad argument");
    parallel_ranges(n, threads > 256 ? 256 : threads, [&](int, int64_t a, int64_t b) {
        for (int64_t k = a; k < b; ++k) h_out[k] = (float)h_in[k];
    });
    return PMF_OK;
}

// out = a / b (b_scalar when h_b == NULL): IEEE float64 division, i.e. exactly NumPy's a / b, by all host cores
extern "C" int pmf_host_divide_f64(const double* h_a, const double* h_b, double b_scalar, int64_t n, double* h_out, int32_t threads) {
    PMF_REQUIRE(n >= 0 && (n == 0 || (h_a && h_out)), "bad argument");
    parallel_ranges(n, threads > 256 ? 256 : threads, [&](int, int64_t lo, int64_t hi) {
        if (h_b) for (int64_t k = lo; k < hi; ++k) h_out[k] = h_a[k] / h_b[k];
        else for (int64_t k = lo; k < hi; ++k) h_out[k] = h_a[k] / b_scalar;
    });
    return PMF_OK;
}
