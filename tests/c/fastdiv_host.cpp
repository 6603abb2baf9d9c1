// Host check of the multiply-high division the kernels decode work items with (csrc/blocks.h,
// FastDiv): every divisor up to 70000 against edge dividends, plus two million random
// (divisor, dividend) pairs below 2^31 -- the bound the table builder enforces per block.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../dtfft_b200/csrc/blocks.h"
using namespace dtfftb;
static unsigned umulhi(unsigned a, unsigned b){ return (unsigned)(((unsigned long long)a*b)>>32); }
static unsigned fdiv(unsigned n, const FastDiv& f){ return f.mul ? (umulhi(n,f.mul)>>f.shr) : n; }
int main(){
  std::mt19937_64 rng(1);
  long long bad=0, checked=0;
  for (unsigned d=1; d<=70000; ++d){
    FastDiv f=FastDiv::make(d);
    unsigned ns[]={0,1,d-1,d,d+1,2*d-1,2*d,0x7fffffffu,0x7ffffffeu,(0x7fffffffu/d)*d,(0x7fffffffu/d)*d-1};
    for(unsigned n:ns){ if(n>0x7fffffffu) continue; checked++; if(fdiv(n,f)!=n/d){bad++; if(bad<5) printf("bad d=%u n=%u got %u want %u\n",d,n,fdiv(n,f),n/d);} }
    for(int k=0;k<20;++k){ unsigned n=(unsigned)(rng()&0x7fffffffu); checked++; if(fdiv(n,f)!=n/d){bad++; if(bad<5) printf("bad d=%u n=%u\n",d,n);} }
  }
  for (int t=0;t<2000000;++t){ unsigned d=(unsigned)(rng()%0x7fffffffu)+1; FastDiv f=FastDiv::make(d); unsigned n=(unsigned)(rng()&0x7fffffffu); checked++; if(fdiv(n,f)!=n/d){bad++; if(bad<5) printf("bad d=%u n=%u\n",d,n);} }
  printf("checked %lld bad %lld\n",checked,bad); return bad!=0;
}
